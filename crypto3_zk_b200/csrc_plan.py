"""Host mirror of ntt_make_plan (csrc/zkb_ntt_plan.h): the radix split of a transform, for work models in bench.py."""

MAX_LOG_R = 8


def ntt_radices(log_n):
    """log2 of the radix of every pass: ceil(log_n / 8) passes, sizes as even as possible, larger ones first."""
    p = max(1, (log_n + MAX_LOG_R - 1) // MAX_LOG_R)
    base, rem = divmod(log_n, p)
    return [base + (1 if i < rem else 0) for i in range(p)]
