#!/usr/bin/env python3
"""Express-path counters of the encode kernel on the headline text (needs the diagnostics build:
scripts/build_variant.sh stats -DXLLM_EXP_STATS, XLLM_INGEST_LIB=build/stats/libxllm_ingest_stats.so)."""
import ctypes, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xllm_service_b200 as x
from xllm_service_b200 import workload, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
natural = len(sys.argv) > 2 and sys.argv[2] == "natural"
L = ctypes.CDLL(_lib.lib_path())
out = (ctypes.c_ulonglong * 16)()
if natural:
    h = x.Ingest(tokenizer_path=os.path.join(ROOT, "tests", "golden", "sp_natural_32k"))
    batch = workload.cut_prompts(workload.natural_corpus(n * 16384), 16384)
    L.xllm_debug_exp_stats(out)
    ids, cnt, st = h.encode_batch(batch.text, batch.offsets, 16384)
    assert (st == 0).all()
else:
    h = x.Ingest(tokenizer_path=os.path.join(ROOT, "tests", "golden", "sp_bpe_8k"))
    vocab = workload.make_vocabulary()
    wb = workload.pack_prompts(vocab)
    _, wcnt, st = h.encode_batch(wb.text, wb.offsets, 32)
    batch, _ = workload.make_prompts_exact_tokens(n, 4096, wcnt, seed=1)
    L.xllm_debug_exp_stats(out)
    ids, cnt, st = h.encode_batch(batch.text, batch.offsets, 4096 + 64)
    assert (st == 0).all() and (cnt == 4096).all()
L.xllm_debug_exp_stats(out)
names = ["attempts", "taken", "forced_drains", "", "fail_not_simple", "fail_not_boundary", "fail_unfinished_word",
         "fail_long_word", "fail_hard_word", "words", "memo_hits", "steps_with_miss", "bytes_done", "bytes_total", "requests_handed_over"]
print(json.dumps({k: int(out[i]) for i, k in enumerate(names) if k}))
