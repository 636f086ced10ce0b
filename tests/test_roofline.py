"""roofline.py keeps the SURVEY 8d work formulas bench.py's roofline fractions are computed from."""
import roofline as R


def test_unit_costs_follow_the_survey():
    assert R.WMAC_PER_FE_MUL == 72 and (R.M_ADD, R.M_MIXED, R.M_DBL) == (12, 11, 8)
    assert R.msm_fixed_wmac(49, 16) == 49 * 16 * 792
    assert R.straus_wmac(2) == (2 * 71 * 792 + 128 * 576)
    assert abs(R.msm_point_wmac(1 << 21, 16) - 1.35e4) / 1.35e4 < 0.05            # "~1.35e4 wMAC/point"
    # SURVEY's estimate with 49 fixed-base terms is 1.7e6 wMAC per verify at w = 16; with the 66 terms the verifier needs: 1.95e6
    assert 1.6e6 < R.verify_wmac(16) - R.msm_fixed_wmac(17, 16) < 1.8e6
    assert R.prove_wmac(16) < 7.0e6
