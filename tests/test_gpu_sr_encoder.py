"""GPU: the super_resolution encoders (trinerflet_b200/sr_encoder.py) through the C-ABI on a B200, against the
reference's own fp32 results (tests/golden/sr_encoder_fp32.npz) and the oracle (oracle/sr_encoder.py; fp32 and fp64 at small
sizes, fp32 at the plane size of the reference's configs (super_resolution/configs/triplane-sr100_400_2.yaml: 16 channels, 1024^2, wavelet scale 16,
low_res_scale 4)).  The same checks run on CPU over the host build of the kernels in tests/test_sr_encoder.py."""
import pytest

from tests import sr_cases

pytestmark = pytest.mark.gpu
dev = "cuda"


def test_wavelet_two_readings_match_reference_golden():
    sr_cases.check_wavelet_golden(dev)


def test_plane_pyramids_match_reference_golden():
    sr_cases.check_kplanes_golden(dev)


def test_position_gradient_properties():
    sr_cases.check_position_gradient_properties(dev)


def test_two_render_step_matches_oracle_small_and_generic_channels():
    sr_cases.check_against_oracle(dev, C=8, R=64, scale=8, low=4, high=1, M=300)
    sr_cases.check_against_oracle(dev, C=24, R=32, scale=2, low=2, high=2, M=100, seed=1)


def test_two_render_step_matches_oracle_config_size():
    """16 channels, 1024^2 planes, four levels, low reading at 256^2 -- the reference's triplane-sr100_400_2 encoder"""
    sr_cases.check_against_oracle(dev, C=16, R=1024, scale=16, low=4, high=1, M=20000, seed=2, exact=False)


def test_low_resolution_phase_runs_coarse_levels_only():
    sr_cases.check_low_resolution_phase_cost(dev, R=512)
