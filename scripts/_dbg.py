import os, sys, json
ROOT="/root/repo"; sys.path.insert(0, ROOT)
import numpy as np
import xllm_service_b200 as x
from xllm_service_b200 import workload
from oracle import oracle as o
model=os.path.join(ROOT,"tests","golden","sp_bpe_8k")
h=x.Ingest(tokenizer_path=model); S=o.SentencePieceOracle(model)
texts=[b"ab ab ab", b"ab "*50, b"hello world", b"the quick brown fox jumps over the lazy dog "*8, b"a b c d e f g h i j k l m n o p q r s t u v w x y z "*6, b"  lead", b"trail  ", b"x"*14+b" yy", b"abcdefghijklmnop qrs"]
vocab=workload.make_vocabulary()
import random
rnd=random.Random(1)
texts.append(b" ".join(rnd.choice(vocab) for _ in range(300)))
b=workload.pack_prompts(texts)
ids,n,st=h.encode_batch(b.text,b.offsets,4096)
for i,t in enumerate(texts):
    exp=S.encode(t).tolist(); got=ids[i,:n[i]].tolist()
    if exp!=got:
        k=next((j for j in range(min(len(exp),len(got))) if exp[j]!=got[j]), min(len(exp),len(got)))
        print("MISMATCH",i,t[:60],"len",len(t),"status",st[i],"first diff at",k,"exp",exp[max(0,k-3):k+6],"got",got[max(0,k-3):k+6],"n",len(exp),len(got))
    else: print("ok",i,len(t))
