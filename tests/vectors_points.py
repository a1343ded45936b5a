"""Shared inputs for the endomorphism tests (CPU host simulation and GPU): G1/G2 points inside and outside the
order-r subgroup with the ground truth [r]P == O from oracle/pyref.py, and scalars that hit the GLV edge cases."""
from oracle import pyref as P


def g1_membership_cases():
    """[(point, in_subgroup)]: members (generator multiples, infinity), random curve points (cofactor component
    present), pure cofactor-subgroup points incl. one of order 3, and member + non-member sums."""
    rng = P.SplitMix64(381)
    h1 = (P.BLS_X - 1) ** 2 // 3
    pts = [P.G1_GEN, None] + [P.g1_mul(rng.fr(), P.G1_GEN) for _ in range(5)]
    for seed in (3, 1000, 2 ** 200 + 17):
        q = P.g1_curve_point(seed)
        t = P._u1(P.curve_mul_unreduced(P.R_MOD, P._w1(q)))          # cofactor subgroup (order divides h1)
        t3 = P._u1(P.curve_mul_unreduced(h1 // 3, P._w1(t)))         # order 3 or infinity
        pts += [q, t, t3, P._u1(P._aff_add(P._w1(t), P._w1(P.G1_GEN)))]
    return [(p, P.g1_in_subgroup(p)) for p in pts]


def g2_membership_cases():
    rng = P.SplitMix64(382)
    pts = [P.G2_GEN, None] + [P.g2_mul(rng.fr(), P.G2_GEN) for _ in range(3)]
    for seed in (5, 77):
        q = P.g2_curve_point(seed)
        t = P.curve_mul_unreduced(P.R_MOD, q)
        pts += [q, t, P._aff_add(t, P.G2_GEN)]
    return [(p, P.g2_in_subgroup(p)) for p in pts]


def glv_scalars():
    L = P.GLV_LAMBDA
    rng = P.SplitMix64(383)
    return [0, 1, 2, 15, 16, L - 1, L, L + 1, 2 * L - 1, 2 * L, L * L - 1, L * L, L * (L + 1), P.R_MOD - 1, P.R_MOD - 2,
            (1 << 128) - 1, 1 << 128, (1 << 254) + 12345, P.R_MOD, P.R_MOD + 5, (1 << 256) - 1] + [rng.fr() for _ in range(12)]
