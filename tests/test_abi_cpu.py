"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares, argument validation works without a GPU, and the host mirror keeps the reference's
class surface and state-dict keys."""
import ctypes as C
import os
import re

import pytest
import torch

import relationnetworks_clevr_b200 as R
from oracle import rn_oracle as O
from relationnetworks_clevr_b200 import _lib
from tests.golden_util import case_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()


def test_header_symbols_are_exported(lib):
    header = open(os.path.join(ROOT, "include", "rn_b200.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*|unsigned long long)\s+(rn_\w+)\s*\(", header, flags=re.M))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.rn_abi_version() == _lib.RN_ABI_VERSION == int(re.search(r"#define RN_ABI_VERSION (\d+)", header).group(1))


def test_struct_mirrors_match_the_compiled_header(lib):
    """The ctypes mirrors of the seven header structs have the sizes the library was compiled with (checked at load time
    too: a stale mirror would be read as garbage fields, e.g. rn_conv_cfg.flags)."""
    mirrors = (_lib.RelationCfg, _lib.FCfg, _lib.ConvCfg, _lib.ConvLayer, _lib.ConvGrads, _lib.LstmCfg, _lib.AdamCfg)
    sizes = (C.c_int32 * 8)()
    assert lib.rn_abi_struct_sizes(sizes, 8) == len(mirrors)
    assert [C.sizeof(m) for m in mirrors] == list(sizes)[:len(mirrors)]
    assert lib.rn_abi_struct_sizes(sizes, 3) == 3 and lib.rn_abi_struct_sizes(None, 3) == 0


def test_argument_validation_without_gpu(lib):
    a, b = C.c_size_t(), C.c_size_t()
    bad = _lib.RelationCfg(4, 64, 26, 128, 255, 4, 0, 0, 1)      # G not a multiple of 4
    assert lib.rn_relation_workspace(C.byref(bad), C.byref(a), C.byref(b)) == -1
    assert b"G must be" in lib.rn_last_error()
    bad = _lib.RelationCfg(4, 64, 26, 128, 256, 4, 4, 0, 1)      # qinj out of range
    assert lib.rn_relation_workspace(C.byref(bad), C.byref(a), C.byref(b)) == -1
    ok = _lib.RelationCfg(4, 64, 26, 128, 256, 4, 0, 0, 1)
    assert lib.rn_relation_workspace(C.byref(ok), C.byref(a), C.byref(b)) == 0
    assert a.value >= 4 * 4 * 4096 * 256 * 4
    sd = _lib.RelationCfg(4, 12, 7, 256, 512, 4, 0, 1, 1)        # SD shape is not a tcgen05 shape
    assert lib.rn_relation_tc_supported(C.byref(sd)) == 0
    assert lib.rn_relation_workspace(C.byref(sd), C.byref(a), C.byref(b)) == -2
    cc = _lib.ConvCfg(4, 100, 1, 1e-5, 0.1)                      # side not a multiple of 16
    assert lib.rn_conv_workspace(C.byref(cc), C.byref(a), C.byref(b)) == -1
    lc = _lib.LstmCfg(4, 20, 83, 32, 256, 1)                     # hidden size 256: no kernel, the host keeps nn.LSTM
    assert lib.rn_lstm_supported(C.byref(lc)) == 0
    assert lib.rn_lstm_workspace(C.byref(lc), C.byref(a), C.byref(b)) == -2
    assert lib.rn_lstm_supported(C.byref(_lib.LstmCfg(4, 20, 83, 32, 128, 1))) == 1
    # NULL pointers are rejected before any CUDA call
    assert lib.rn_relation_fwd(C.byref(ok), None, None, None, None, None, None, None, None) == -1


class _Args:
    qdict_size, adict_size = 82, 28


@pytest.mark.parametrize("config", ["original-fp", "ir-fp", "original-sd", "ir-sd"])
def test_state_dict_keys_match_reference(config):
    hyp = O.HYPERPARAMS[config]
    m = R.RN(_Args, hyp)
    keys = {k: tuple(v.shape) for k, v in m.state_dict().items() if "num_batches_tracked" not in k}
    want = dict(O.param_shapes(hyp, 82, 28))
    want.update(O.buffer_shapes())
    assert keys == want


@pytest.mark.parametrize("stem", ["ckpt_original_fp", "ckpt_ir_fp"])
def test_shipped_checkpoints_load(stem):
    hyp, p = case_params(stem)
    m = R.RN(_Args, hyp)
    res = m.load_state_dict(p, strict=False)
    assert not res.unexpected_keys
    assert all("num_batches_tracked" in k for k in res.missing_keys)
    # a DataParallel-style wrapper prefixes "module." exactly like the reference's checkpoints
    wrapped = torch.nn.Sequential()
    wrapped.add_module("module", m)
    assert all(k.startswith("module.") for k in wrapped.state_dict())


def test_cpu_inputs_fail_loudly():
    m = R.RN(_Args, O.HYPERPARAMS["original-fp"])
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 3, 128, 128), torch.zeros(1, 5, dtype=torch.long))
    with pytest.raises(RuntimeError, match="CUDA"):
        m.rl.relation(torch.zeros(1, 64, 26), torch.zeros(1, 128))


def test_config_json_matches_reference_hyperparams():
    import json
    cfg = json.load(open(os.path.join(ROOT, "config.json")))["hyperparams"]
    assert cfg == O.HYPERPARAMS
