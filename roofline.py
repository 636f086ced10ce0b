"""Algorithmic integer work of the hot path, the figure `bench.py`'s roofline fractions are computed from.

Unit: one 32x32->64 multiply-accumulate (IMAD.WIDE.U32), "wMAC".  A modular multiplication on 8 x 32-bit limbs is the
64-wMAC schoolbook product plus 8 wMAC for the reduction fold (high half x 977; the x 2^32 part is a shift).  Squarings
count as multiplications; additions, subtractions, small multiples and carry handling (ALU pipe) are not counted.
Group operations are counted with the complete projective formulas for a = 0 (Renes-Costello-Batina 2016): addition
12 M, mixed addition 11 M, doubling 8 M -- the kernels' own XYZZ / Jacobian formulas and the GLV split execute fewer
multiplications than this count, so fractions quoted on it are "reference-algorithm work per second", not pipe occupancy;
the hardware-side number is ncu's sm__pipe_fmaheavy_cycles_active for the same launch (profiles/r1_ncu_full_final_*).
"""
WMAC_PER_FE_MUL = 72
M_MIXED, M_ADD, M_DBL = 11, 12, 8
FIXED_TERMS_VERIFY = (17, 49)                         # pt = ps_tau g + <g_vec, pn_tau>; final commit over g, g_vec, h_vec
FIXED_TERMS_PROVE = 466                               # 2 + 17 + 7 + 38 + 21 + 42 + 43 + 4 x (49 + 25)
VAR_GROUPS_VERIFY = (5, 2, 2, 2, 2)                   # transcript-separated joint ladders over the 13 proof points
NORMALISATIONS_VERIFY, NORMALISATIONS_PROVE = 5, 18


def windows(window_bits: int) -> int:
    """windows of the fixed-base tables: unsigned digits up to 20 bits, signed digits (one spare bit for the carry) above"""
    w = abs(window_bits)
    return (257 + w - 1) // w if (window_bits < 0 or window_bits > 20) else (256 + w - 1) // w


def msm_fixed_wmac(terms: int, window_bits: int) -> float:
    """sum of `terms` fixed-base scalar multiplications from window tables: one mixed addition per (term, window)."""
    return terms * windows(window_bits) * M_MIXED * WMAC_PER_FE_MUL


def straus_wmac(npoints: int) -> float:
    """joint ladder over `npoints` variable points, 4-bit windows with GLV halves: 128 shared doublings plus,
    per point, 7 table additions and 64 window additions, all counted at the mixed-addition cost."""
    adds, dbls = npoints * (7 + 64), 128
    return (adds * M_MIXED + dbls * M_DBL) * WMAC_PER_FE_MUL


def normalisation_wmac(batch: int = 8) -> float:
    """one batched affine normalisation: 3 M per element plus the shared 270 M inversion over `batch` elements."""
    return WMAC_PER_FE_MUL * (3 + 270.0 / batch)


def verify_wmac(window_bits: int, sec1_decode: bool = False) -> float:
    """One u64 verify.  SURVEY 8d estimated 49 fixed-base terms; the verifier needs 66: the 17-term `pt` (circuit.rs:206)
    enters the first WNLA commitment, which the transcript absorbs before the 49-term base case can be formed.
    sec1_decode adds the 14 square roots (253 S + 13 M each) of decompressing a 525-byte record."""
    w = (sum(msm_fixed_wmac(t, window_bits) for t in FIXED_TERMS_VERIFY) + sum(straus_wmac(p) for p in VAR_GROUPS_VERIFY)
         + NORMALISATIONS_VERIFY * normalisation_wmac())
    return w + (14 * 266 * WMAC_PER_FE_MUL if sec1_decode else 0)


def prove_wmac(window_bits: int) -> float:
    return msm_fixed_wmac(FIXED_TERMS_PROVE, window_bits) + 3 * straus_wmac(2) + NORMALISATIONS_PROVE * normalisation_wmac()


def msm_point_wmac(n: int, c: int) -> float:
    """Pippenger with signed c-bit windows: ceil(256/c) mixed additions per point plus the amortised bucket reduction."""
    nwin = (256 + c - 1) // c
    return (nwin + (2 ** c) * nwin / max(n, 1)) * M_MIXED * WMAC_PER_FE_MUL


if __name__ == "__main__":
    for w in (16, 20):
        print(f"W={w}: verify {verify_wmac(w):.3e} wMAC, prove {prove_wmac(w):.3e} wMAC")
    print(f"MSM n=2^21 c=16: {msm_point_wmac(1 << 21, 16):.3e} wMAC/point")
