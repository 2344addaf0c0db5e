"""Generates tests/golden/encoder_query.npz by running the UNMODIFIED reference class
/root/reference/densephrases/encoder.py:Encoder (fp32, CPU, eager) on seeded random weights
(densephrases_b200.encoder.random_state_dict) and synthetic token batches.  Only inputs, seeds and outputs are stored;
the weights are regenerated from the seed (torch CPU generator, same image on the GPU box).
Run in the build container (needs /root/reference):  python tests/golden/make_encoder_golden.py"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from densephrases_b200.encoder import BertGeometry, random_state_dict, synthetic_query_batch  # noqa: E402
from oracle import encoder_ref  # noqa: E402


def load_reference_encoder():
    spec = importlib.util.spec_from_file_location('ref_encoder', '/root/reference/densephrases/encoder.py')
    mod = importlib.util.module_from_spec(spec)
    sys.modules['ref_encoder'] = mod
    spec.loader.exec_module(mod)
    return mod.Encoder


if __name__ == '__main__':
    from transformers import BertConfig, BertModel
    torch.manual_seed(0)
    seed, vocab = 20240923, 28996
    geo = BertGeometry(vocab_size=vocab)
    cfg = BertConfig(vocab_size=vocab, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                     max_position_embeddings=512, type_vocab_size=2, layer_norm_eps=1e-12, hidden_act='gelu',
                     hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)
    RefEncoder = load_reference_encoder()
    model = RefEncoder(cfg, tokenizer=None, transformer_cls=BertModel).eval()
    sd = random_state_dict(geo, seed)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(not m.startswith(('query_start_encoder.encoder', 'query_start_encoder.embeddings.word', 'query_end_encoder.encoder')) for m in missing), missing
    out = {}
    for name, (B, S) in {'b4_s64': (4, 64), 'b3_s24': (3, 24), 'b2_s100': (2, 100)}.items():
        ids, mask, tt = synthetic_query_batch(B, S, vocab, seed + S)
        with torch.no_grad():
            qs, qe = model(input_ids_=ids, attention_mask_=mask, token_type_ids_=tt, return_query=True)
        rs, re_ = encoder_ref.embed_query(sd, ids, mask, tt)
        d = max((qs - rs).abs().max().item(), (qe - re_).abs().max().item())
        print(name, 'reference class vs torch restatement: max abs diff', d, '| out scale', qs.abs().mean().item())
        assert d < 2e-4, d
        out[f'{name}_ids'], out[f'{name}_mask'], out[f'{name}_tt'] = ids.numpy(), mask.numpy(), tt.numpy()
        out[f'{name}_start'], out[f'{name}_end'] = qs.numpy(), qe.numpy()
    out['seed'] = np.array(seed)
    out['vocab'] = np.array(vocab)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'encoder_query.npz'), **out)
    print('wrote tests/golden/encoder_query.npz')
