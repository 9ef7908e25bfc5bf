"""BERT WordPiece tokenisation, bit-exact against the `tokenizers` library.

The reference tokenises with HuggingFace `AutoTokenizer` / `BertTokenizerFast` (polus/models.py:275-284, call shape in
tests/test_models.py:18-25: max_length, padding="max_length", truncation=True, token_type_ids + attention_mask).  This
module restates that pipeline -- the Rust `tokenizers` components BertNormalizer, BertPreTokenizer, WordPiece,
BertProcessing, LongestFirst truncation, fixed-length padding -- in plain Python, so the input side of the hot path has
no TensorFlow / transformers dependency; tests/test_tokenization_cpu.py checks ids, token_type_ids and attention_mask
against `tokenizers` itself on random vocabularies and Unicode text.

Not covered: character offset mappings, and "added tokens" other than the five BERT specials.
"""
import unicodedata

import numpy as np

SPECIALS = ("[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]")
# Unicode White_Space property (Rust char::is_whitespace)
_WHITE_SPACE = frozenset([0x09, 0x0A, 0x0B, 0x0C, 0x0D, 0x20, 0x85, 0xA0, 0x1680, 0x2028, 0x2029, 0x202F, 0x205F, 0x3000]
                         + list(range(0x2000, 0x200B)))


def _is_whitespace(c):
    return ord(c) in _WHITE_SPACE


def _is_control(c):
    if c in "\t\n\r":
        return False
    return unicodedata.category(c).startswith("C")  # Rust: general category "Other" (Cc, Cf, Cn, Co, Cs)


def _is_chinese_char(cp):
    return (0x4E00 <= cp <= 0x9FFF or 0x3400 <= cp <= 0x4DBF or 0x20000 <= cp <= 0x2A6DF or 0x2A700 <= cp <= 0x2B73F
            or 0x2B740 <= cp <= 0x2B81F or 0x2B920 <= cp <= 0x2CEAF or 0xF900 <= cp <= 0xFAFF or 0x2F800 <= cp <= 0x2FA1F)


def _is_punctuation(c):
    cp = ord(c)
    if 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126:
        return True
    return unicodedata.category(c).startswith("P")


def bert_normalize(text, clean_text=True, handle_chinese_chars=True, strip_accents=None, lowercase=True):
    """tokenizers.normalizers.BertNormalizer: clean -> CJK spacing -> strip accents -> lowercase."""
    if clean_text:
        out = []
        for c in text:
            if c == "\0" or c == "�" or _is_control(c):
                continue
            out.append(" " if _is_whitespace(c) else c)
        text = "".join(out)
    if handle_chinese_chars:
        out = []
        for c in text:
            if _is_chinese_char(ord(c)):
                out.extend((" ", c, " "))
            else:
                out.append(c)
        text = "".join(out)
    if strip_accents is None:
        strip_accents = lowercase
    if strip_accents:
        text = "".join(c for c in unicodedata.normalize("NFD", text) if unicodedata.category(c) != "Mn")
    if lowercase:
        text = text.lower()
    return text


def bert_pre_tokenize(text):
    """tokenizers.pre_tokenizers.BertPreTokenizer: split on whitespace (dropped), isolate every punctuation character."""
    words, cur = [], []
    for c in text:
        if _is_whitespace(c):
            if cur:
                words.append("".join(cur))
                cur = []
        elif _is_punctuation(c):
            if cur:
                words.append("".join(cur))
                cur = []
            words.append(c)
        else:
            cur.append(c)
    if cur:
        words.append("".join(cur))
    return words


class BertWordPieceTokenizer:
    """vocab: dict token -> id, an iterable of tokens in id order, or the path of a vocab.txt."""

    def __init__(self, vocab, lowercase=True, strip_accents=None, handle_chinese_chars=True, clean_text=True,
                 unk_token="[UNK]", sep_token="[SEP]", cls_token="[CLS]", pad_token="[PAD]", mask_token="[MASK]",
                 wordpieces_prefix="##", max_input_chars_per_word=100):
        if isinstance(vocab, str):
            with open(vocab, encoding="utf-8") as f:
                vocab = [line.rstrip("\n") for line in f]
        if not isinstance(vocab, dict):
            vocab = {t: i for i, t in enumerate(vocab)}
        self.vocab = dict(vocab)
        self.ids_to_tokens = {i: t for t, i in self.vocab.items()}
        self.lowercase, self.strip_accents = lowercase, strip_accents
        self.handle_chinese_chars, self.clean_text = handle_chinese_chars, clean_text
        self.unk_token, self.sep_token, self.cls_token, self.pad_token, self.mask_token = unk_token, sep_token, cls_token, pad_token, mask_token
        self.prefix = wordpieces_prefix
        self.max_chars = max_input_chars_per_word
        for t in (unk_token, sep_token, cls_token, pad_token):
            if t not in self.vocab:
                raise ValueError(f"special token {t!r} is not in the vocabulary")
        self._specials = [t for t in (unk_token, sep_token, cls_token, pad_token, mask_token) if t in self.vocab]

    # ------------------------------------------------------------------ pieces
    def _wordpiece(self, word):
        if len(word) > self.max_chars:
            return [self.unk_token]
        pieces, start = [], 0
        while start < len(word):
            end, cur = len(word), None
            while start < end:
                sub = word[start:end]
                if start > 0:
                    sub = self.prefix + sub
                if sub in self.vocab:
                    cur = sub
                    break
                end -= 1
            if cur is None:
                return [self.unk_token]
            pieces.append(cur)
            start = end
        return pieces

    def _split_specials(self, text):
        """Special tokens written literally in the text are matched before normalisation (the library's added-vocabulary
        step): earliest match first, the longest token among those starting at the same position."""
        parts, pos = [], 0
        while pos < len(text):
            best, best_at = None, len(text)
            for t in self._specials:
                at = text.find(t, pos)
                if at != -1 and (at < best_at or (at == best_at and len(t) > len(best))):
                    best, best_at = t, at
            if best is None:
                break
            if best_at > pos:
                parts.append((text[pos:best_at], False))
            parts.append((best, True))
            pos = best_at + len(best)
        if pos < len(text):
            parts.append((text[pos:], False))
        return parts

    def tokenize(self, text):
        out = []
        for part, is_special in self._split_specials(text):
            if is_special:
                out.append(part)
                continue
            norm = bert_normalize(part, self.clean_text, self.handle_chinese_chars, self.strip_accents, self.lowercase)
            for word in bert_pre_tokenize(norm):
                out.extend(self._wordpiece(word))
        return out

    def convert_tokens_to_ids(self, tokens):
        unk = self.vocab[self.unk_token]
        return [self.vocab.get(t, unk) for t in tokens]

    # ------------------------------------------------------------------ encoding
    def encode(self, text, text_pair=None, max_length=None, padding=False, truncation=False, add_special_tokens=True):
        a = self.convert_tokens_to_ids(self.tokenize(text))
        b = self.convert_tokens_to_ids(self.tokenize(text_pair)) if text_pair is not None else None
        n_special = (2 if b is None else 3) if add_special_tokens else 0
        if truncation and max_length is not None:
            total = len(a) + (len(b) if b is not None else 0) + n_special
            to_remove = total - max_length
            if to_remove > 0:
                na, nb = len(a), (len(b) if b is not None else 0)
                if b is None:
                    na = max(0, na - to_remove)
                else:
                    # tokenizers' LongestFirst (utils/truncation.rs): with n1 the shorter length and L the room left after
                    # the special tokens, only the longer sequence is cut (to L - n1) when that suffices; otherwise
                    # both are cut to L/2, and the longer one (the second on ties) keeps the odd token
                    room = max_length - n_special
                    n1, n2, swap = na, nb, False
                    if n1 > n2:
                        n1, n2, swap = n2, n1, True
                    n2 = n1 if n1 > room else max(n1, room - n1)
                    if n1 + n2 > room:
                        n1 = room // 2
                        n2 = n1 + room % 2
                    if swap:
                        n1, n2 = n2, n1
                    na, nb = max(n1, 0), max(n2, 0)
                a = a[:na]
                if b is not None:
                    b = b[:nb]
        cls, sep, pad = self.vocab[self.cls_token], self.vocab[self.sep_token], self.vocab[self.pad_token]
        if add_special_tokens:
            ids = [cls] + a + [sep]
            types = [0] * len(ids)
            if b is not None:
                ids += b + [sep]
                types += [1] * (len(b) + 1)
        else:
            ids = a + (b or [])
            types = [0] * len(a) + [1] * len(b or [])
        mask = [1] * len(ids)
        if padding in (True, "max_length") and max_length is not None and len(ids) < max_length:
            n = max_length - len(ids)
            ids += [pad] * n
            types += [0] * n
            mask += [0] * n
        return {"input_ids": ids, "token_type_ids": types, "attention_mask": mask}

    def __call__(self, text, text_pair=None, max_length=None, padding=False, truncation=False, return_tensors=None,
                 return_token_type_ids=True, return_attention_mask=True, **unused):
        """Same call shape as the HuggingFace tokenizers the reference uses; batches -> int32 numpy arrays when every row
        has the same length (padding="max_length"), otherwise lists."""
        single = isinstance(text, str)
        texts = [text] if single else list(text)
        pairs = [text_pair] if (single and text_pair is not None) else (list(text_pair) if text_pair is not None else [None] * len(texts))
        rows = [self.encode(t, p, max_length=max_length, padding=padding, truncation=truncation) for t, p in zip(texts, pairs)]
        keys = ["input_ids"] + (["token_type_ids"] if return_token_type_ids else []) + (["attention_mask"] if return_attention_mask else [])
        out = {}
        for k in keys:
            col = [r[k] for r in rows]
            if single and return_tensors is None:
                out[k] = col[0]
            elif len({len(c) for c in col}) == 1:
                out[k] = np.asarray(col, dtype=np.int32)
            else:
                out[k] = col
        return out
