"""one flow extra of bench.py on its own: python profiles/quick_flows.py groth16|placeholder [log]"""
import sys, os, json, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from crypto3_zk_b200 import Context
what = sys.argv[1]
log = int(sys.argv[2]) if len(sys.argv) > 2 else (22 if what == "groth16" else 20)
args = argparse.Namespace(groth16_log=log, placeholder_log=log)
ctx = Context(0)
dev = torch.device("cuda", 0)
fn = {"groth16": bench.groth16_extra, "placeholder": bench.placeholder_extra, "placeholder_prover": bench.placeholder_prover_extra}[what]
print(json.dumps(fn(args, torch, ctx, dev), indent=1))
