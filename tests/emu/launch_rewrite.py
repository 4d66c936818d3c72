"""Test infrastructure: rewrites CUDA triple-chevron launches into emulator launches (see tests/emu/host_emu.h).

    name<<<grid, block, smem, stream>>>(args);   ->   emu::launch(dim3(grid), dim3(block), [&] { name(args); });
"""
import re

_HEAD = re.compile(r"([A-Za-z_][\w:]*(?:<[^<>;(){}]*>)?)\s*<<<")


def _balanced(text, start, open_ch, close_ch):
    """index just past the bracket that closes the one at text[start]"""
    depth = 0
    for i in range(start, len(text)):
        if text[i] == open_ch:
            depth += 1
        elif text[i] == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
    raise ValueError("unbalanced")


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite_launches(src: str) -> str:
    out, pos = "", 0
    while True:
        m = _HEAD.search(src, pos)
        if not m:
            return out + src[pos:]
        cfg_end = src.index(">>>", m.end())
        cfg = _split_top(src[m.end():cfg_end])
        paren = src.index("(", cfg_end)
        args_end = _balanced(src, paren, "(", ")")
        semi = src.index(";", args_end)
        out += src[pos:m.start()]
        out += f"emu::launch(dim3({cfg[0]}), dim3({cfg[1]}), [&] {{ {m.group(1)}{src[paren:args_end]}; }});"
        pos = semi + 1
