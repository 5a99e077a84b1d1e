#!/usr/bin/env python3
"""Encode kernels alone on natural text (this image's site-packages sources / docs cut into 16 KB prompts,
SentencePiece BPE 32 000 fixture): python scripts/bench_natural.py [n_prompts] — device-resident, CUDA events,
bit-exact gate of a few prompts against the CPU oracle."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xllm_service_b200 as x
from xllm_service_b200 import workload
from oracle import oracle as o
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
check = int(sys.argv[2]) if len(sys.argv) > 2 else 16
d_sp = os.path.join(ROOT, "tests", "golden", "sp_natural_32k")
h = x.Ingest(tokenizer_path=d_sp)
pb = workload.cut_prompts(workload.natural_corpus(n * 16384), 16384)
n = pb.n
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
d_text = torch.from_numpy(pb.text).cuda(); d_off = torch.from_numpy(pb.offsets).cuda()
d_ids = torch.empty((n, 16384), dtype=torch.int32, device="cuda")
d_n = torch.empty(n, dtype=torch.int32, device="cuda"); d_st = torch.empty(n, dtype=torch.int32, device="cuda")
run = lambda: h.encode_batch_device(n, d_text.data_ptr(), d_off.data_ptr(), d_ids.data_ptr(), 16384, d_n.data_ptr(), d_st.data_ptr(), stream.cuda_stream)
run(); torch.cuda.synchronize()
st = d_st.cpu().numpy(); nid = d_n.cpu().numpy()
assert (st == 0).all(), np.unique(st)
S = o.SentencePieceOracle(d_sp)
ids = d_ids[:check].cpu().numpy()
bad = 0
for r in range(check):
    want = S.encode(pb.text[pb.offsets[r]:pb.offsets[r + 1]].tobytes()).tolist()
    got = ids[r, :nid[r]].tolist()
    if got != want:
        bad += 1
        k = next((j for j in range(min(len(want), len(got))) if want[j] != got[j]), min(len(want), len(got)))
        print("MISMATCH prompt", r, "at id", k, "want", want[max(0, k - 3):k + 5], "got", got[max(0, k - 3):k + 5], len(want), len(got), file=sys.stderr)
ts = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); run(); e1.record(stream); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[1]
print(json.dumps({"natural_sp32k_ms": ms, "prompts": int(n), "MB_per_s": pb.text.size / ms / 1e3, "tokens": int(nid.sum()), "oracle_mismatches": bad, "checked": check}))
