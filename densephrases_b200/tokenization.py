"""WordPiece tokenisation of questions into the `[CLS] q [SEP] [PAD]...` features the query towers consume
(reference: squad_utils.py question-only path :117-141,401-431,594-608 via HF BertTokenizer.encode_plus, padded to
max_query_length, token_type_ids all 0).  CPU work, kept (not accelerated), written from the published BERT algorithm:
text cleaning + whitespace / CJK / punctuation basic tokenisation, then greedy longest-match-first WordPiece; pinned token for
token against transformers.BertTokenizer on a shared vocabulary (tests/test_api.py::test_tokenizer_matches_transformers_bert_tokenizer).
No vocabulary file is reachable offline; `WordPieceTokenizer.from_pretrained_or_synthetic` loads `vocab.txt` when a
directory has one and otherwise builds a deterministic synthetic cased vocabulary of the SpanBERT size (28 996)."""
import os
import unicodedata

PAD, UNK, CLS, SEP, MASK = '[PAD]', '[UNK]', '[CLS]', '[SEP]', '[MASK]'
_SPECIAL_IDS = {PAD: 0, UNK: 100, CLS: 101, SEP: 102, MASK: 103}


def _is_punct(ch):
    cp = ord(ch)
    if (33 <= cp <= 47) or (58 <= cp <= 64) or (91 <= cp <= 96) or (123 <= cp <= 126):
        return True
    return unicodedata.category(ch).startswith('P')


def _is_whitespace(ch):
    return ch in ' \t\n\r' or unicodedata.category(ch) == 'Zs'


def _is_control(ch):
    return ch not in '\t\n\r' and unicodedata.category(ch).startswith('C')


def _is_cjk(cp):
    return ((0x4E00 <= cp <= 0x9FFF) or (0x3400 <= cp <= 0x4DBF) or (0x20000 <= cp <= 0x2A6DF) or (0x2A700 <= cp <= 0x2B73F) or
            (0x2B740 <= cp <= 0x2B81F) or (0x2B820 <= cp <= 0x2CEAF) or (0xF900 <= cp <= 0xFAFF) or (0x2F800 <= cp <= 0x2FA1F))


def basic_tokenize(text, do_lower_case=False):
    """BERT BasicTokenizer: drop NUL / U+FFFD / control characters, every Unicode space -> ' ', CJK ideographs become words of
    their own, whitespace split, (lower-case + strip accents), split at punctuation."""
    cleaned = []
    for ch in text:
        cp = ord(ch)
        if cp == 0 or cp == 0xFFFD or _is_control(ch):
            continue
        if _is_whitespace(ch):
            cleaned.append(' ')
        elif _is_cjk(cp):
            cleaned.append(' ' + ch + ' ')
        else:
            cleaned.append(ch)
    out = []
    for tok in ''.join(cleaned).split():
        if do_lower_case:
            tok = ''.join(c for c in unicodedata.normalize('NFD', tok.lower()) if unicodedata.category(c) != 'Mn')
        cur = ''
        for ch in tok:
            if _is_punct(ch):
                if cur:
                    out.append(cur)
                    cur = ''
                out.append(ch)
            else:
                cur += ch
        if cur:
            out.append(cur)
    return out


class WordPieceTokenizer(object):
    def __init__(self, vocab, do_lower_case=False, max_chars=100):
        self.vocab = vocab
        self.ids_to_tokens = {i: t for t, i in vocab.items()}
        self.do_lower_case = do_lower_case
        self.max_chars = max_chars
        self.pad_token_id, self.unk_token_id = vocab[PAD], vocab[UNK]
        self.cls_token_id, self.sep_token_id = vocab[CLS], vocab[SEP]

    @classmethod
    def from_vocab_file(cls, vocab_file, do_lower_case=False):
        vocab = {line.rstrip('\n'): i for i, line in enumerate(open(vocab_file, encoding='utf-8'))}
        return cls(vocab, do_lower_case)

    @classmethod
    def from_pretrained_or_synthetic(cls, path=None, do_lower_case=False, vocab_size=28996, extra_words=()):
        vocab_file = os.path.join(path, 'vocab.txt') if path and os.path.isdir(path) else None
        if vocab_file and os.path.exists(vocab_file):
            return cls.from_vocab_file(vocab_file, do_lower_case)
        vocab = {f'[unused{i}]': i for i in range(vocab_size)}          # placeholder rows, overwritten below
        vocab = {}
        for tok, i in _SPECIAL_IDS.items():
            vocab[tok] = i
        nxt = 104
        chars = [chr(c) for c in range(33, 127)]
        for piece in chars + ['##' + c for c in chars] + sorted(set(extra_words)):
            while nxt in _SPECIAL_IDS.values():
                nxt += 1
            if piece not in vocab and nxt < vocab_size:
                vocab[piece] = nxt
                nxt += 1
        return cls(vocab, do_lower_case)

    def wordpiece(self, word):
        if len(word) > self.max_chars:
            return [UNK]
        pieces, start = [], 0
        while start < len(word):
            end, cur = len(word), None
            while start < end:
                sub = word[start:end] if start == 0 else '##' + word[start:end]
                if sub in self.vocab:
                    cur = sub
                    break
                end -= 1
            if cur is None:
                return [UNK]
            pieces.append(cur)
            start = end
        return pieces

    def tokenize(self, text):
        return [p for w in basic_tokenize(text, self.do_lower_case) for p in self.wordpiece(w)]

    def encode_question(self, text, max_query_length=64):
        """-> (input_ids, attention_mask, token_type_ids, tokens): [CLS] + pieces (truncated) + [SEP], zero padded."""
        toks = [CLS] + self.tokenize(text)[:max_query_length - 2] + [SEP]
        ids = [self.vocab.get(t, self.unk_token_id) for t in toks]
        pad = max_query_length - len(ids)
        return ids + [self.pad_token_id] * pad, [1] * len(ids) + [0] * pad, [0] * max_query_length, toks
