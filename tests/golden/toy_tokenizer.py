"""Deterministic toy tokenizer shared by gen_train_goldens.py and the CPU test: `<|...|>` special tokens get ids from
100000 upwards in first-seen order of a FIXED vocabulary list, every other character is its code point."""
import re

SPECIALS = (["<|begin_of_text|>", "<|start_header_id|>", "<|end_header_id|>", "<|eot_id|>"] +
            [f"<|reserved_special_token_{i}|>" for i in range(251)])


class _Enc:
    def __init__(self, ids):
        self.input_ids = ids


class ToyTokenizer:
    bos_token_id = 100000

    def __init__(self):
        self.table = {s: 100000 + i for i, s in enumerate(SPECIALS)}
        self.pat = re.compile(r"<\|[^|]+\|>")

    def __call__(self, text, add_special_tokens=False):
        ids, pos = [], 0
        for m in self.pat.finditer(text):
            ids += [ord(c) for c in text[pos:m.start()]]
            ids.append(self.table[m.group(0)])
            pos = m.end()
        ids += [ord(c) for c in text[pos:]]
        if add_special_tokens:
            ids = [self.bos_token_id] + ids
        return _Enc(ids)
