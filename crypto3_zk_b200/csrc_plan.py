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
