"""Encoder timing: B questions x S tokens through both towers (22.05 GFLOP per question at S=64)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from densephrases_b200.encoder import BertGeometry, Encoder, random_state_dict, synthetic_query_batch
B, S = int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 64
geo = BertGeometry()
enc = Encoder(geo, state_dict=random_state_dict(geo, 1))
ids, mask, tt = (t.cuda() for t in synthetic_query_batch(B, S, geo.vocab_size, 2))
mode = sys.argv[3] if len(sys.argv) > 3 else "tf32"
enc.set_precision({"precise": "3xtf32"}.get(mode, mode))     # tf32 | 3xtf32 | bf16x3
n_iter = int(sys.argv[4]) if len(sys.argv) > 4 else 20
for _ in range(3): enc.embed_query(ids, mask, tt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = n_iter
e0.record()
for _ in range(n): enc.embed_query(ids, mask, tt)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
T = B * S
flops = 2 * 12 * (2.0 * T * 768 * (2304 + 768 + 3072 + 3072) + 2.0 * B * 12 * S * S * 64 * 2)
print(f"B={B} S={S} mode={enc.mode}: {ms:.3f} ms/forward  {B/ms*1000:.0f} questions/s  {flops/ms/1e9:.1f} TFLOP/s (dense tf32 peak ~ half of bf16)")
