"""Golden vectors for the truecaser, produced by the UNMODIFIED reference class: the source of `class TrueCaser`
(/root/reference/densephrases/utils/squad_utils.py:1452-1585) is cut out of the file with `ast` and executed as is (the module itself
imports h5py / old transformers APIs that are absent here; the class needs only os, pickle, math, string and whitespace_tokenize, the
latter taken the same way from utils/data_utils.py).  A small synthetic distribution file is generated from a seeded corpus.
Run in the build container:  python tests/golden/make_truecase_golden.py  ->  tests/golden/truecase.dist, tests/golden/truecase.json"""
import ast
import json
import math
import os
import pickle
import random
import string
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))


def cut(path, name):
    src = open(path).read()
    node = next(n for n in ast.parse(src).body if getattr(n, "name", None) == name)
    return ast.get_source_segment(src, node)


ns = {"os": os, "pickle": pickle, "math": math, "string": string}
exec(cut("/root/reference/densephrases/utils/data_utils.py", "whitespace_tokenize"), ns)
exec(cut("/root/reference/densephrases/utils/squad_utils.py", "TrueCaser"), ns)

rng = random.Random(7)
words = ["paris", "Paris", "PARIS", "us", "US", "Us", "apple", "Apple", "who", "Who", "is", "the", "The", "president", "President", "of", "france", "France",
         "new", "New", "york", "York", "in", "did", "Did", "obama", "Obama", "win", "what", "What", "nba", "NBA", "may", "May", "turkey", "Turkey", "bill", "Bill",
         "gates", "Gates", "'s", "river", "seine", "Seine", "?", ",", "1999", "a", "A", "an", "mr.", "Mr.", "it", "IT", "It"]
corpus = []
for _ in range(900):
    n = rng.randint(3, 9)
    corpus.append([rng.choice(words) for _ in range(n)])
uni, bwd, fwd, tri, lookup = defaultdict(int), defaultdict(int), defaultdict(int), defaultdict(int), defaultdict(set)
for sent in corpus:
    for i, w in enumerate(sent):
        uni[w] += 1
        lookup[w.lower()].add(w)
        if i > 0:
            bwd[sent[i - 1] + "_" + w] += 1
        if i + 1 < len(sent):
            fwd[w + "_" + sent[i + 1].lower()] += 1
        if 0 < i < len(sent) - 1:
            tri[sent[i - 1] + "_" + w + "_" + sent[i + 1].lower()] += 1
dist = {"uni_dist": dict(uni), "backward_bi_dist": dict(bwd), "forward_bi_dist": dict(fwd), "trigram_dist": dict(tri),
        "word_casing_lookup": {k: sorted(v) for k, v in lookup.items()}}       # lists: a fixed iteration order travels with the pickle
pickle.dump(dist, open(os.path.join(HERE, "truecase.dist"), "wb"), protocol=2)

# the reference indexes the count tables with [] -> give it defaultdicts over the same counts
ref_tables = {k: (defaultdict(int, v) if k != "word_casing_lookup" else v) for k, v in dist.items()}
pickle.dump(ref_tables, open("/tmp/truecase_ref.dist", "wb"))
ref = ns["TrueCaser"]("/tmp/truecase_ref.dist")
sentences = ["who is the president of france ?", "what is the us", "did obama win in new york", "bill gates 's apple", "may", "", "  the   river seine , 1999 ",
             "zzz unknown-word in paris", "it is a turkey", "mr. gates", "?", "nba"]
for _ in range(300):
    sentences.append(" ".join(rng.choice(words + ["qqq", "o'neil", "12", "x-ray"]).lower() for _ in range(rng.randint(1, 10))))
out = [{"sentence": s, "oov": o, "truecased": ref.get_true_case(s, o)} for s in sentences for o in ("title", "lower", "as-is")]
scores = []
for _ in range(200):
    tok = rng.choice([w for w in words if len(lookup[w.lower()]) > 1])
    prev, nxt = rng.choice([None] + words), rng.choice([None] + words)
    scores.append({"prev": prev, "token": tok, "next": nxt, "score": ref.get_score(prev, tok, nxt)})
json.dump({"cases": out, "scores": scores}, open(os.path.join(HERE, "truecase.json"), "w"), ensure_ascii=True, indent=0)
print(len(out), len(scores))
