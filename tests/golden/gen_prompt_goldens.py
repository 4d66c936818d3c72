"""Golden prompts produced by EXECUTING the reference's own prompt-building lines
(evaluations/models/llava_multiimg_siglip_anyres.py:412-442, sliced out of `run_llava_local_inference` and run with the
module's `head` / `tail` constants and eval_utils' `retain_n_images` / `get_instruction`).  Authoring container only
(needs /root/reference).  Writes tests/golden/ref_prompts.json."""
import ast
import json
import re
import textwrap
import types
from pathlib import Path

REF = Path("/root/reference/evaluations/models")
OUT = Path(__file__).resolve().parent

CASES = [  # (question, setting, ques_type, image_num, num_patchs_per_images_real)
    ("<image><image><image><image> What is shown on slide 3?", "direct", "open", 4, [6, 6, 6, 6]),
    ("What is the title?", "direct", "open", 1, [0]),
    ("Compare <image> and <image>.", "cot", "multiple-choice", 2, [16, 4]),
    ("<image> only one tag but three images", "none", "open", 3, [2, 0, 9]),
    ("no tags, two images", "direct", "multiple-choice", 2, [4, 4]),
    ("<image><image><image> three tags, two images", "direct", "open", 2, [1, 2]),
    ("<image><image><image><image><image> five tags, two images", "direct", "open", 2, [1, 2]),
    ("<image> budget exhausted: no tiles list", "direct", "open", 1, []),
    ("tabs \r\n\t\t\r\n\t\t inside <image>", "cot", "open", 1, [3]),
]


def load_reference():
    src = (REF / "llava_multiimg_siglip_anyres.py").read_text()
    lines = src.splitlines()
    body = textwrap.dedent("\n".join(lines[411:442]))                      # lines 412-442
    consts = "\n".join(l for l in lines[:30] if l.startswith("head =") or l.startswith("tail ="))
    tree = ast.parse((REF / "eval_utils.py").read_text())
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("retain_n_images", "get_instruction")]
    mod = types.ModuleType("ref_prompt")
    mod.__dict__["re"] = re
    exec(compile(ast.Module(body=fns, type_ignores=[]), "eval_utils.py", "exec"), mod.__dict__)
    exec(consts, mod.__dict__)
    return mod, compile(body, "anyres_412_442", "exec")


def main():
    mod, code = load_reference()
    out = []
    for question, setting, ques_type, image_num, tiles in CASES:
        env = dict(mod.__dict__)
        env.update(question=question, image_num=image_num, num_patchs_per_images_real=list(tiles),
                   instruction=mod.get_instruction(setting, ques_type))
        rec = {"question": question, "setting": setting, "ques_type": ques_type, "image_num": image_num, "tiles": tiles}
        try:
            exec(code, env)
            rec["prompt"] = env["prompt"]
        except Exception as e:  # noqa: BLE001  (the reference's own failure mode is part of the behaviour)
            rec["raises"] = type(e).__name__
        out.append(rec)
        print(rec.get("raises") or rec["prompt"][:100].replace("\n", "\\n"))
    (OUT / "ref_prompts.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
