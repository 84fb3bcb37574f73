"""K3 / K4 (tcgen05 implicit-GEMM conv) against torch fp32 convolutions on the same 16-bit operands.
Tolerances are storage precision: fp16 raw outputs 1e-3 (11-bit mantissa), bf16 activated outputs 8e-3 (8-bit
mantissa), fp32 after_conv output 1e-4 — all relative to the tensor's max."""
import pytest

pytestmark = pytest.mark.gpu
TOL = {"raw": 1e-3, "praw": 1e-3, "act": 8e-3, "pact": 8e-3, "feat": 1e-4}


def _cases():
    import conv_cases
    return sorted(conv_cases.CASES)


@pytest.mark.parametrize("name", _cases())
def test_conv_case(name):
    import conv_cases
    res = conv_cases.run_case(name, **conv_cases.CASES[name])
    for k, v in res.items():
        if k == "raw_untouched":
            assert v, "conv wrote outside its channel slice"
        else:
            assert v <= TOL[k], (name, k, v)
