"""A/B: encoder forward (B=64, S=64) under the three GEMM schedules of dph_gemm_tf32_set_mode (0 tile per CTA, 1 multicast clusters,
2 persistent 128x256 tiles)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from densephrases_b200 import _lib as L
from densephrases_b200.encoder import BertGeometry, Encoder, random_state_dict, synthetic_query_batch
geo = BertGeometry()
enc = Encoder(geo, state_dict=random_state_dict(geo, 1))
ids, mask, tt = (t.cuda() for t in synthetic_query_batch(64, 64, geo.vocab_size, 2))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for on in (0, 2, 1, 0, 2):
    L.check(L.lib().dph_gemm_tf32_set_mode(on))
    for _ in range(3): enc.embed_query(ids, mask, tt)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20): enc.embed_query(ids, mask, tt)
    e1.record(); torch.cuda.synchronize()
    print(f"gemm mode {on}: {e0.elapsed_time(e1) / 20:.3f} ms/forward", flush=True)
