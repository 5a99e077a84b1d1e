#!/usr/bin/env python3
"""Micro-benchmark of the SentencePiece-BPE encode kernel alone (text resident in HBM)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xllm_service_b200 as x  # noqa: E402
from xllm_service_b200 import workload  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=8192)
ap.add_argument("--T", type=int, default=4096)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--check", type=int, default=4)
a = ap.parse_args()

model = os.path.join(ROOT, "tests", "golden", "sp_bpe_8k")
h = x.Ingest(tokenizer_path=model)
vocab = workload.make_vocabulary()
t0 = time.time()
wb = workload.pack_prompts(vocab)
_, wcnt, st = h.encode_batch(wb.text, wb.offsets, 32)
assert (st == 0).all()
batch, _ = workload.make_prompts_exact_tokens(a.n, a.T, wcnt, seed=1)
print("gen %.1fs, text bytes/prompt %.0f" % (time.time() - t0, batch.text.size / a.n), file=sys.stderr)

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
d_text = torch.from_numpy(batch.text).cuda()
d_off = torch.from_numpy(batch.offsets).cuda()
stride = a.T + 64
d_ids = torch.zeros((a.n, stride), dtype=torch.int32, device="cuda")
d_n = torch.zeros(a.n, dtype=torch.int32, device="cuda")
d_st = torch.zeros(a.n, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()


def run():
    h.encode_batch_device(a.n, d_text.data_ptr(), d_off.data_ptr(), d_ids.data_ptr(), stride, d_n.data_ptr(),
                          d_st.data_ptr(), stream.cuda_stream)


for _ in range(a.warmup):
    run()
torch.cuda.synchronize()
ts = []
for _ in range(a.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
assert (d_n.cpu().numpy() == a.T).all(), d_n.cpu().numpy()[:10]
assert (d_st.cpu().numpy() == 0).all()
if a.check:
    from oracle import oracle as o
    S = o.SentencePieceOracle(model)
    ids = d_ids[:a.check].cpu().numpy()
    for i in range(a.check):
        assert ids[i, :a.T].tolist() == S.encode(batch.prompt(i)).tolist()
ms = sorted(ts)[len(ts) // 2]
byts = batch.text.size + 4 * a.n * a.T
print(json.dumps({"kernel": "sp_encode", "n": a.n, "T": a.T, "ms_median": ms, "ms_min": min(ts),
                  "algo_GBps": byts / ms / 1e6, "frac_of_6585": byts / ms / 1e6 / 6585.1,
                  "req_per_s": a.n / ms * 1e3, "text_bytes_per_prompt": batch.text.size / a.n}))
