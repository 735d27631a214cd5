"""CPU oracle for the crypto3-zk hot paths (NTT/LDE/FRI-commit and MSM).

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (crypto3_zk_b200/) may
import, call, link or execute anything in this package.  Allowed users are
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg.

What it restates: the *algorithms* crypto3-zk calls for its two data-parallel
hot paths.  The arithmetic itself lives in un-vendored sibling libraries
(crypto3-algebra, crypto3-math, crypto3-hash, crypto3-containers; effective
pin = umbrella NilFoundation/crypto3@1bd56b12f410f3f1a4891076705a9261a6b1efaa,
/root/reference/.github/workflows/pull-request.yml:29), so the restatement
follows their published algorithms (libff/libfqfft lineage) and is anchored on
the reference's own call sites, tests and literal vectors.

Parity pinning status (see DESIGN.md "Oracle"):
  * G1/G2 MSM + Fr arithmetic, BLS12-381: PINNED by the bellperson-generated
    literals in test/systems/ppzksnark/r1cs_gg_ppzksnark/
    r1cs_gg_ppzksnark_aggregation_conformity.cpp (tests/golden/bls12_381_ipp2.json).
  * Keccak-256 transcript: PINNED by test/transcript/transcript.cpp:50-64.
  * KZG commit: PINNED algebraically (commit == f(alpha)*G, test/commitment/kzg.cpp:97).
  * NTT / LDE / Merkle roots: the reference holds no concrete vector for these
    ("parity unpinned" by reference fixtures); pinned only structurally
    (test/commitment/fri.cpp:122-123, fold_polynomial.cpp:52-135) and by the
    DFT definition itself.  omega and layout conventions live in one table
    (oracle/fields.py, oracle/fri.py) so they can be corrected in one place.
"""
