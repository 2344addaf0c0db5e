"""Statistical truecaser for lower-cased questions -- the `TrueCaser` of the reference (densephrases/utils/squad_utils.py:1452-1585;
used at model.py:52,66-67 and utils/open_utils.py:147-154): per token, choose among the casings seen in training the one with the
best product of unigram, backward-bigram, forward-bigram and trigram relative frequencies (pseudo-count 5), reading the pickled
n-gram counts of `truecase/english_with_questions.dist` (options.py:83).

Same file format, same API (`TrueCaser(path).get_true_case(sentence, out_of_vocabulary_token_option)`), same outputs (pinned by
tests/golden/truecase.json, generated from the reference class); the scoring is restated as a sum of four log-ratios."""
import math
import pickle
import string

_PSEUDO = 5.0
_KEYS = ("uni_dist", "backward_bi_dist", "forward_bi_dist", "trigram_dist", "word_casing_lookup")


class TrueCaser(object):
    def __init__(self, dist_file_path):
        with open(dist_file_path, "rb") as f:
            tables = pickle.load(f)
        missing = [k for k in _KEYS if k not in tables]
        if missing:
            raise KeyError(f"{dist_file_path}: not a truecaser distribution file (missing {missing})")
        self.uni_dist, self.backward_bi_dist, self.forward_bi_dist, self.trigram_dist, self.word_casing_lookup = (tables[k] for k in _KEYS)

    @staticmethod
    def _count(table, key):
        # the reference indexes nltk FreqDist / defaultdict objects (a missing key counts 0 and, for defaultdict, gets inserted);
        # .get gives the same value for both and for plain dicts without mutating the tables
        v = table.get(key, 0)
        return v if v is not None else 0

    def _log_ratio(self, table, make_key, candidate, casings):
        num = self._count(table, make_key(candidate)) + _PSEUDO
        den = 0
        for alt in casings:
            den += self._count(table, make_key(alt)) + _PSEUDO
        return math.log(num / den)

    def get_score(self, prev_token, possible_token, next_token):
        casings = self.word_casing_lookup[possible_token.lower()]
        score = self._log_ratio(self.uni_dist, lambda t: t, possible_token, casings)
        # the reference adds the four logs left to right; missing contexts contribute log(1) = 0.0 exactly
        score += self._log_ratio(self.backward_bi_dist, lambda t: prev_token + "_" + t, possible_token, casings) if prev_token is not None else 0.0
        if next_token is not None:
            next_token = next_token.lower()
            score += self._log_ratio(self.forward_bi_dist, lambda t: t + "_" + next_token, possible_token, casings)
        else:
            score += 0.0
        if prev_token is not None and next_token is not None:
            score += self._log_ratio(self.trigram_dist, lambda t: prev_token + "_" + t + "_" + next_token, possible_token, casings)
        else:
            score += 0.0
        return score

    @staticmethod
    def first_token_case(raw):
        return raw[:1].upper() + raw[1:]

    def get_true_case(self, sentence, out_of_vocabulary_token_option="title"):
        tokens = sentence.strip().split()                    # whitespace_tokenize (utils/data_utils.py)
        out = []
        for i, token in enumerate(tokens):
            if token in string.punctuation or token.isdigit():
                out.append(token)
                continue
            token = token.lower()
            casings = self.word_casing_lookup.get(token) if hasattr(self.word_casing_lookup, "get") else self.word_casing_lookup[token]
            if casings:
                if len(casings) == 1:
                    out.append(next(iter(casings)))
                else:
                    prev_token = out[i - 1] if i > 0 else None
                    next_token = tokens[i + 1] if i + 1 < len(tokens) else None
                    best, best_score = None, float("-inf")
                    for cand in casings:                     # first strictly better candidate wins, in the set's iteration order
                        s = self.get_score(prev_token, cand, next_token)
                        if s > best_score:
                            best, best_score = cand, s
                    out.append(best)
                if i == 0:
                    out[0] = self.first_token_case(out[0])
            elif out_of_vocabulary_token_option == "title":
                out.append(token.title())
            else:                                            # "lower" and "as-is" coincide: the token was lower-cased above
                out.append(token)
        return "".join(t if (t.startswith("'") or t in string.punctuation) else " " + t for t in out).strip()


def truecase_questions(truecaser, questions):
    """model.py:66-67 / open_utils.py:154: only all-lower-case questions are re-cased."""
    return [truecaser.get_true_case(q) if q == q.lower() else q for q in questions]
