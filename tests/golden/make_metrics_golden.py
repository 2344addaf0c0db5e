"""Golden vectors for the answer-matching metrics, produced by the UNMODIFIED reference functions
(/root/reference/densephrases/utils/eval_utils.py:9-86, loaded by file path; its only non-stdlib import, ujson, is aliased to json).
Run in the build container:  python tests/golden/make_metrics_golden.py  ->  tests/golden/metrics.json"""
import importlib.util
import json
import os
import random
import sys

sys.modules.setdefault("ujson", json)
spec = importlib.util.spec_from_file_location("ref_eval_utils", "/root/reference/densephrases/utils/eval_utils.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

rng = random.Random(11)
vocab = ["the", "The", "a", "an", "river", "Seine", "Paris,", "1999", "yes", "no", "noanswer", "U.S.", "café", "café", "New-York", "of", "treaty",
         "(1648)", "Westphalia", "“quoted”", "it's", "AN", "A", "  ", "\t", "x"]


def phrase():
    return " ".join(rng.choice(vocab) for _ in range(rng.randint(0, 6)))


pairs = [("the Seine", "Seine"), ("yes", "no"), ("", ""), ("A river.", "a  river"), ("noanswer", "the noanswer")] + [(phrase(), phrase()) for _ in range(400)]
patterns = [r"Sein(e)?", r"19\d\d", r"(unclosed", r"^the\s+river", r"\bparis\b", r"café", r".*treaty"]
out = {"pairs": [], "regex": []}
for p, g in pairs:
    f1 = ref.f1_score(p, g)
    out["pairs"].append({"prediction": p, "truth": g, "norm_p": ref.normalize_answer(p), "f1": [float(v) for v in f1],
                         "em": bool(ref.exact_match_score(p, g)), "drqa_em": bool(ref.drqa_exact_match_score(p, g)),
                         "drqa_norm": ref.drqa_normalize(p)})
for p, _ in pairs[:120]:
    for pat in patterns:
        out["regex"].append({"prediction": p, "pattern": pat, "match": bool(ref.drqa_regex_match_score(p, pat))})
out["max_over"] = [{"prediction": p, "truths": [g, p.upper(), "zzz"],
                    "em": bool(ref.drqa_metric_max_over_ground_truths(ref.drqa_exact_match_score, p, [g, p.upper(), "zzz"]))} for p, g in pairs[:60]]
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "metrics.json"), "w"), ensure_ascii=True, indent=0)
print(len(out["pairs"]), len(out["regex"]))
