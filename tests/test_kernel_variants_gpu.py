"""GPU parity of the opt-in / fallback kernel variants of the tcgen05 relation path.

The variant switches (RN_B200_*) are read once per process inside librn_b200.so, so each variant runs the
same oracle comparison as tests/test_parity_gpu.py::test_relation_tcgen05_matches_oracle in a fresh interpreter.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = {
    "generator_warpgroup_everywhere": {"RN_B200_GENWG": "1"},
    "classic_form_everywhere": {"RN_B200_GENWG": "0"},
    "stored_dz4_h1_two_pass_dgrad": {"RN_B200_REGEN_DZ4": "0", "RN_B200_REGEN_H1": "0", "RN_B200_DGRAD_PASSES": "2"},
    "boundary_covering_job_order": {"RN_B200_SCHED": "1"},
    "text_encoder_on_main_stream": {"RN_B200_TEXT_STREAM": "0"},
}

SCRIPT = r"""
import sys
sys.path.insert(0, {root!r})
import tests.test_parity_gpu as T
for name in T.TC_CASES:
    T.test_relation_tcgen05_matches_oracle(name, "parity")
stem = [c for c in T.CASES if "d12" not in c][0]
T.test_model_eval_matches_reference_golden(stem, "auto")
T.test_model_train_step_matches_reference_golden(stem, "auto")
T.test_relation_tcgen05_grid_sweep(144)
print("VARIANT_OK")
"""


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_variant_matches_oracle(variant):
    env = dict(os.environ)
    env.update(VARIANTS[variant])
    out = subprocess.run([sys.executable, "-c", SCRIPT.format(root=ROOT)], env=env, cwd=ROOT, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0 and "VARIANT_OK" in out.stdout, (out.stdout[-2000:], out.stderr[-2000:])
