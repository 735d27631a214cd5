"""Host mirror of ntt_make_plan (csrc/zkb_ntt_plan.h): the radix split of a transform, for work models in bench.py."""

MAX_LOG_R = 8


def ntt_radices(log_n, small_first=False):
    """log2 of the radix of every pass: ceil(log_n / 8) passes, sizes as even as possible; the larger ones first, or
    last for the zero-padded forward transform of an LDE (small_first)."""
    p = max(1, (log_n + MAX_LOG_R - 1) // MAX_LOG_R)
    base, rem = divmod(log_n, p)
    if small_first:
        return [base + (1 if i >= p - rem else 0) for i in range(p)]
    return [base + (1 if i < rem else 0) for i in range(p)]


def ntt_pass_products(log_r, zero_levels=0):
    """Field products per element of one ntt_pass_kernel tile transform (csrc/zkb_ntt_pass.cuh): the level loop of
    ZKB_NTT_FOR_EACH_PHASE with the unit twiddles of block 0 skipped (radix-2 phase: b * c only for blk != 0; radix-4
    phase: 4 products per 4 elements, 1 when blk == 0), without the inter-pass / coset table products."""
    lh = log_r - 1 - zero_levels
    total = 0.0
    if (lh + 1) & 1:
        lvl = log_r - 1 - lh
        total += 0.5 * (1.0 - 2.0 ** -lvl)
        lh -= 1
    while lh >= 1:
        lvl = log_r - 1 - lh
        first = 2.0 ** -lvl                 # share of the tasks that sit in block 0
        total += (1.0 - first) * 1.0 + first * 0.25
        lh -= 2
    return total


def ntt_products_per_element(log_n, small_first=False, zero_levels=0, known_log=0, coset=False, inverse=False):
    """Products per OUTPUT element of one transform as ntt_device_t runs it: tile transforms + one inter-pass twiddle
    per pass boundary (1/N rides on T_1 or on the coset table) + the coset scale; with known outputs (LDE) every pass
    after the first works on 1 - 2^-known_log of the tiles and the first pass does not scale the rows it never stores."""
    lr = ntt_radices(log_n, small_first)
    p = len(lr)
    known = known_log >= 3 and p >= 2 and known_log + 3 <= lr[0]
    keep = 1.0 - 2.0 ** -known_log if known else 1.0
    total = 0.0
    for i, r in enumerate(lr):
        per = ntt_pass_products(r, zero_levels if i == 0 else 0)
        if i + 1 < p:
            per += keep if i == 0 else 1.0          # inter-pass twiddle at the store
        total += per * (keep if i > 0 else 1.0)
    if p == 1 and inverse and not coset:
        total += 1.0                                # scalar 1/N store table
    if coset:
        total += 1.0 if not known else 1.0         # load (forward) or store (inverse) table
    return total
