"""Host-side prompt construction of the reference's evaluation loop — pure string logic.

`build_eval_prompt` restates evaluations/models/llava_multiimg_siglip_anyres.py:412-442: the Llama-3 chat frame
(`head` / `tail`, :22-23), the `<image>` bookkeeping and the per-image placeholder text
`image k: <|reserved_special_token_20|>` + (1 + tiles) x `<|reserved_special_token_195|>` + `<|reserved_special_token_21|>`.
Every `<|reserved_special_token_195|>` (id 128200 = LeopardConfig.image_token_index) becomes ONE crop = 169 decoder rows
in the merge.  Behaviour is kept as it is in the reference, including its two quirks:
  * the placeholder loop runs over the `<image>` count of the ORIGINAL question, so `<image>` tags the loop prepends for
    images the question did not mention stay literal text;
  * with more `<image>` tags than images, `retain_n_images(question, count - image_num)` keeps count - image_num tags (not
    image_num), and the loop indexes the crop list by tag number (IndexError when it runs past the images).
Pinned by tests/golden/ref_prompts.json, produced by executing those reference lines (tests/golden/gen_prompt_goldens.py).
"""
from __future__ import annotations

import re
from typing import Sequence

HEAD = "<|begin_of_text|><|start_header_id|>user<|end_header_id|>\n\n"       # EV/...anyres.py:22
TAIL = "<|eot_id|><|start_header_id|>assistant<|end_header_id|>\n\n"          # :23
DEFAULT_IMAGE_TOKEN = "<image>"
IMG_OPEN, IMG_SLOT, IMG_CLOSE = ("<|reserved_special_token_20|>", "<|reserved_special_token_195|>",
                                 "<|reserved_special_token_21|>")


def get_instruction(setting: str, ques_type: str) -> str:
    """evaluations/models/eval_utils.py:104-119."""
    table = {
        True: {"cot": "First think step by step. Then answer with the letter of the correct option.",
               "direct": "Answer with the option's letter from the given choices directly.", "none": ""},
        False: {"cot": "First think step by step. Then answer with a single word or phrase.",
                "direct": "Answer the question using a single word or phrase.", "none": ""},
    }
    return table[ques_type == "multiple-choice"][setting]


def retain_n_images(text: str, n: int) -> str:
    """Keep the first n `<image>` tags, drop the rest (evaluations/models/eval_utils.py:122-146)."""
    seen = 0

    def keep(m):
        nonlocal seen
        seen += 1
        return m.group(0) if seen <= n else ""
    return re.sub(r"<image>", keep, text)


def image_placeholder(k: int, n_crops: int) -> str:
    """Text that stands for image number k (1-based) with `n_crops` ViT inputs (original + tiles), :434-435."""
    return f"image {k}: {IMG_OPEN}{IMG_SLOT * n_crops}{IMG_CLOSE}"


def build_eval_prompt(question: str, instruction: str, image_num: int,
                      num_patchs_per_images_real: Sequence[int]) -> str:
    """The prompt string the reference tokenises (…anyres.py:412-442).  `num_patchs_per_images_real[i]` = tiles of image
    i (crops - 1), as computed at :394 / :401."""
    n_tags = question.count(DEFAULT_IMAGE_TOKEN)
    if n_tags < image_num:
        question = f"{DEFAULT_IMAGE_TOKEN * (image_num - n_tags)} {question}"
    elif n_tags > image_num:
        question = retain_n_images(question, n_tags - image_num)
    prompt = f"{HEAD}{question}\n{instruction}{TAIL}"
    for k in range(1, n_tags + 1):                    # the ORIGINAL tag count (reference quirk, see module docstring)
        crops = 1
        if num_patchs_per_images_real:
            crops += num_patchs_per_images_real[k - 1]
        prompt = prompt.replace(DEFAULT_IMAGE_TOKEN, image_placeholder(k, crops), 1)
    return prompt.replace("\r\n\t\t\r\n\t\t", " ")
