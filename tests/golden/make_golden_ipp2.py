#!/usr/bin/env python3
"""Extract the literal BLS12-381 known-answer vectors the reference's own tests hold for the MSM
path into tests/golden/bls12_381_ipp2.json.

Source (read-only, only in the build container):
  /root/reference/test/systems/ppzksnark/r1cs_gg_ppzksnark/r1cs_gg_ppzksnark_aggregation_conformity.cpp
    :578-862   bls381_polynomial_test        (Fr products: coefficients + product-form evaluation)
    :864-930   bls381_prove_commitment_test  (G1 and G2 MSMs of size 16 over SRS powers)
    :1065-1884 bls381_gipa_tipp_mipp_test    (G1 MSMs of size 4,2,1: z_c pairs, final_c)
Only literals (inputs and expected outputs) are extracted; no code is copied.
Run once in the build container:  python tests/golden/make_golden_ipp2.py
"""
import json
import os
import re

SRC = ("/root/reference/test/systems/ppzksnark/r1cs_gg_ppzksnark/"
       "r1cs_gg_ppzksnark_aggregation_conformity.cpp")
HEX = re.compile(r"0x([0-9a-fA-F]+)_cppui_modular(\d+)")


def hexes(lines, lo, hi, bits=None):
    """All literals on 1-based lines [lo, hi], optionally only those with the given width tag."""
    out = []
    for ln in lines[lo - 1:hi]:
        for h, b in HEX.findall(ln):
            if bits is None or int(b) == bits:
                out.append(int(h, 16))
    return out


def find(lines, needle, start=1):
    for i in range(start - 1, len(lines)):
        if needle in lines[i]:
            return i + 1
    raise KeyError(needle)


def main():
    lines = open(SRC).read().split("\n")
    g = {"source": SRC.replace("/root/reference/", ""), "curve": "bls12_381"}

    # ---- bls381_polynomial_test
    t0 = find(lines, "BOOST_AUTO_TEST_CASE(bls381_polynomial_test)")
    t1 = find(lines, "BOOST_AUTO_TEST_CASE(bls381_prove_commitment_test)")
    fr = hexes(lines, t0, t1 - 1, 255)
    # r_shift, 8 tr, 256 coeffs, kzg_challenge, eval
    assert len(fr) == 1 + 8 + 256 + 2, len(fr)
    g["polynomial"] = {"r_shift": fr[0], "tr": fr[1:9], "coeffs": fr[9:265],
                       "kzg_challenge": fr[265], "eval": fr[266]}

    # ---- bls381_prove_commitment_test
    t2 = find(lines, "BOOST_AUTO_TEST_CASE(bls381_transcript_test)")
    fr = hexes(lines, t1, t2 - 1, 255)
    fq = hexes(lines, t1, t2 - 1, 381)
    assert len(fr) == 7 and len(fq) == 12, (len(fr), len(fq))
    g["prove_commitment"] = {
        "n": 8, "alpha": fr[0], "beta": fr[1], "kzg_challenge": fr[2], "tr": fr[3:6], "r_shift": fr[6],
        # G2 points as [[x.c0, x.c1], [y.c0, y.c1]]
        "comm_v": [[[fq[0], fq[1]], [fq[2], fq[3]]], [[fq[4], fq[5]], [fq[6], fq[7]]]],
        "comm_w": [[fq[8], fq[9]], [fq[10], fq[11]]],
    }

    # ---- bls381_gipa_tipp_mipp_test: c, r, ch, ch_inv, gp_z_c, gp_final_c
    t3 = find(lines, "BOOST_AUTO_TEST_CASE(bls381_gipa_tipp_mipp_test)")
    t4 = find(lines, "BOOST_AUTO_TEST_CASE(bls381_prove_tipp_mipp_test)")
    lc = find(lines, "constexpr std::array<G1_value_type, n> c = {", t3)
    lr = find(lines, "constexpr std::array<scalar_field_value_type, n> r = {", lc)
    c = hexes(lines, lc, lr - 1, 381)
    assert len(c) == 16
    lg = find(lines, "gipa_tipp_mipp<curve_type>(", lr)
    r = hexes(lines, lr, lg - 1, 255)
    assert len(r) == 8
    lch = find(lines, "std::vector<scalar_field_value_type> ch = {", lg)
    lci = find(lines, "std::vector<scalar_field_value_type> ch_inv = {", lch)
    ch = hexes(lines, lch, lci - 1, 255)
    lend = find(lines, "};", lci)
    ch_inv = hexes(lines, lci, lend, 255)
    assert len(ch) == 3 and len(ch_inv) == 3
    lz = find(lines, "gp_z_c = {", lci)
    lf = find(lines, "G1_value_type gp_final_c = G1_value_type(", lz)
    lze = find(lines, "    };", lz)
    z = hexes(lines, lz, lze, 381)
    assert len(z) == 12, len(z)
    lfe = find(lines, ");", lf)
    fc = hexes(lines, lf, lfe + 3, 381)[:2]
    assert t3 < lc < lfe < t4
    g["gipa"] = {
        "c": [[c[2 * i], c[2 * i + 1]] for i in range(8)], "r": r, "ch": ch, "ch_inv": ch_inv,
        "z_c": [[[z[4 * i], z[4 * i + 1]], [z[4 * i + 2], z[4 * i + 3]]] for i in range(3)],
        "final_c": fc,
    }

    def enc(o):
        if isinstance(o, int) and not isinstance(o, bool) and o > 1 << 31:
            return hex(o)
        if isinstance(o, list):
            return [enc(x) for x in o]
        if isinstance(o, dict):
            return {k: enc(v) for k, v in o.items()}
        return o

    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bls12_381_ipp2.json")
    with open(out, "w") as f:
        json.dump(enc(g), f, indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
