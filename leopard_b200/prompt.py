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


# ----------------------------------------------------------------------------------------------------------------
# training twin (Pai-Megatron-Patch/megatron_patch/data/llava/mm_pretrain_dataset.py:769-874, `preprocess_llama3`)
# ----------------------------------------------------------------------------------------------------------------
MEGATRON_IMAGE_TOKEN_INDEX = -200          # data/llava/constants.py:8; one per FEATURE ROW (169 per crop)
TRAIN_IMAGE_SLOT = "<|reserved_special_token_250|>"
_PROMPT_ROLES = ("玩家", "system", "策略", "领域知识", "长期记忆", "user")   # turns that carry no loss (:836)


def train_image_placeholder(k: int, per_image_token: int, tiles, slot: str = TRAIN_IMAGE_SLOT) -> str:
    """Placeholder text of image k in a training sample: the crops arrive as tiles + [global] (…:247-253), every crop
    is `per_image_token` slot tokens, tile t is closed by the tag `<|reserved_special_token_{22+t}|>`, the global view
    by `<|reserved_special_token_21|>` (:801-809).  `tiles` None = no any-res plan (one crop)."""
    text = f"image {k}: {IMG_OPEN}"
    if tiles is not None:
        for t in range(tiles):
            text += slot * per_image_token + f"<|reserved_special_token_{22 + t}|>"
    return text + slot * per_image_token + IMG_CLOSE


def build_train_sample(sources, tokenizer, num_image: int = 1, per_image_token: int = 576,
                       num_patches_per_image: Sequence[int] = (), img_special_token: str = TRAIN_IMAGE_SLOT):
    """Token ids / labels / loss weights / preference tags of one Llama-3 chat sample, as `preprocess_llama3` builds
    them: `<image>` tags are stripped from each turn and their placeholders prepended to it (images numbered across
    turns), missing tags are prepended to the first turn, a sample with MORE tags than images is dropped (None); every
    slot token becomes -200; only assistant message bodies (incl. `<|eot_id|>`) carry weight 1 / preference 2.
    `tokenizer(text, add_special_tokens=False).input_ids` and `.bos_token_id` are all that is used of the tokenizer.
    Returns dict(input_ids, labels, weights, preferences) of [n_samples, L] tensors (torch.stack, like the reference)."""
    import torch
    role_of = {"human": "user", "gpt": "assistant"}
    plans = list(num_patches_per_image)
    slot_id = tokenizer(img_special_token, add_special_tokens=False).input_ids[0]
    rows = {"input_ids": [], "labels": [], "weights": [], "preferences": []}
    for turns in sources:
        if "from" in sources[0] and turns[0]["from"] != "human":
            turns = turns[1:]                                  # a leading non-human turn is skipped (:785-787)
        text_of = lambda t: t["content"] if "content" in t else t["value"]  # noqa: E731
        n_tags = sum(text_of(t).count(DEFAULT_IMAGE_TOKEN) for t in turns)
        if num_image > n_tags:
            turns[0]["content"] = DEFAULT_IMAGE_TOKEN * (num_image - n_tags) + turns[0]["content"]
        elif num_image < n_tags:
            return None
        ids, weights, prefs = [tokenizer.bos_token_id], [0.0], [0]
        k = 1
        for t in turns:
            role = role_of[t["from"]] if "from" in t else t["role"]
            raw = t["value"] if "value" in t else t["content"]
            prefix = ""
            for _ in range(raw.count(DEFAULT_IMAGE_TOKEN)):
                prefix += train_image_placeholder(k, per_image_token, plans[k - 1] if plans else None, img_special_token)
                k += 1
            body = prefix + raw.replace(DEFAULT_IMAGE_TOKEN, "")
            head = tokenizer(f"<|start_header_id|>{role}<|end_header_id|>\n\n", add_special_tokens=False).input_ids
            tail = tokenizer(f"{body}<|eot_id|>", add_special_tokens=False).input_ids
            ids += head + tail
            learn = role not in _PROMPT_ROLES
            weights += [0.0] * len(head) + [1.0 if learn else 0.0] * len(tail)
            prefs += [0] * len(head) + [2 if learn else 0] * len(tail)
        rows["labels"].append(torch.tensor(ids, dtype=torch.long))       # labels keep the slot id (shifted later)
        rows["input_ids"].append(torch.tensor([MEGATRON_IMAGE_TOKEN_INDEX if i == slot_id else i for i in ids],
                                              dtype=torch.long))
        rows["weights"].append(torch.tensor(weights, dtype=torch.float))
        rows["preferences"].append(torch.tensor(prefs, dtype=torch.long))
    return {name: torch.stack(v, dim=0) for name, v in rows.items()}
