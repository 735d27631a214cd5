#!/usr/bin/env python3
"""profiles/ntt_traffic.json from an ncu_summary.py text of one LDE (16 x 2^20 -> 2^23): python profiles/make_ntt_traffic.py SUMMARY.txt"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
path = sys.argv[1]
total, launches = 0.0, 0
unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
for block in open(path).read().split("---")[1:]:
    launches += 1
    for line in block.splitlines():
        p = line.split()
        if len(p) >= 3 and p[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            total += float(p[1].replace(",", "")) * unit[p[2]]
out = {"source": "%s (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of the %d launches - ntt_pass_kernel and ntt_known_scatter_kernel - "
                 "of one LDE of 16 Pallas-Fq polynomials 2^20 -> 2^23)" % (path, launches),
       "dram_bytes_per_polynomial": total / 16, "launches_per_capture": launches, "polynomials_per_capture": 16,
       "note": "bench.py scales this to its own launch count: traffic per launch = dram_bytes_per_polynomial * polys_per_step / launches_per_step",
       "ntt_sources_sha256": bench.ntt_sources_hash()}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ntt_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
