"""The reference's evaluation loop (evaluations/models/llava_multiimg_siglip_anyres.py:364-500, `run_llava_local_inference`)
with leopard_b200 under it — what a maintainer's script looks like after the switch described in INTEGRATION.md.

Reference lines                              here
  :373-375  from_pretrained / eval / to      hf.myLlavaForConditionalGeneration (same calls; synthetic weights with --synthetic)
  :385-405  PIL crop pipeline + processor    leopard_b200.frontend.ImageFrontend (GPU, bit-identical pixel_values)
  :407-446  prompt with image placeholders   leopard_b200.prompt.build_eval_prompt gives the exact prompt STRING; the
                                             tokenizer is not available offline, so build_input_ids() below emits the ids
                                             that string tokenises to around the images (195 placeholder id per crop)
  :448-454  llava.generate + slice           unchanged
Run:  python examples/llava_eval_loop.py --synthetic --layers 2          (needs a B200; no CPU fallback)
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from leopard_b200 import hf  # noqa: E402
from leopard_b200.config import leopard_llava, reduced  # noqa: E402
from leopard_b200.frontend import ImageFrontend  # noqa: E402
from leopard_b200.weights import SyntheticCheckpoint  # noqa: E402


def build_input_ids(cfg, question_ids, crops_per_image):
    """`image k: <20> <195> x (1 + tiles) <21>` per image in front of the question (EV/...anyres.py:426-440): every
    <|reserved_special_token_195|> (cfg.image_token_index) is one crop = 169 decoder rows after the merge."""
    ids = []
    for n in crops_per_image:
        ids += [cfg.image_token_index - 175, *([cfg.image_token_index] * n), cfg.image_token_index - 174]
    return torch.tensor([ids + list(question_ids)], dtype=torch.long)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--checkpoint", help="HF-layout Leopard-LLaVA directory (config.json + safetensors)")
    ap.add_argument("--synthetic", action="store_true", help="seeded random weights (no checkpoint files offline)")
    ap.add_argument("--layers", type=int, default=0, help="with --synthetic: reduced ViT/decoder depth for a quick run")
    ap.add_argument("--max-new-tokens", type=int, default=16)
    args = ap.parse_args()
    if args.synthetic:
        cfg = leopard_llava()
        if args.layers:
            cfg = reduced(cfg, vit_layers=args.layers, dec_layers=args.layers, vocab=cfg.dec.vocab)
        ck = SyntheticCheckpoint(cfg, seed=1234)
        llava = hf.myLlavaForConditionalGeneration.from_state_dict(cfg, lambda n: ck[n])
    else:
        llava = hf.myLlavaForConditionalGeneration.from_pretrained(args.checkpoint, torch_dtype=torch.float32)
        cfg = llava.config
    llava.eval()
    llava.to("cuda:0")
    frontend = ImageFrontend(llava.device)

    rng = np.random.default_rng(0)
    examples = [  # (images as RGB uint8 arrays — Image.open(p).convert('RGB') in the reference —, question token ids)
        ([rng.integers(0, 256, (1152, 896, 3), dtype=np.uint8) for _ in range(4)], rng.integers(0, 1000, 24)),
        ([rng.integers(0, 256, (384, 384, 3), dtype=np.uint8)], rng.integers(0, 1000, 8)),
    ]
    with torch.no_grad():
        for images, question in examples:
            pixel_values, crops_per_image = frontend(images)                    # [N, 3, 364, 364] on the GPU
            input_ids = build_input_ids(cfg, question.tolist(), crops_per_image)
            input_token_len = input_ids.shape[1]
            output_ids = llava.generate(input_ids, pixel_values=pixel_values, attention_mask=torch.ones_like(input_ids),
                                        pad_token_id=0, eos_token_id=[128001, 128009],
                                        max_new_tokens=args.max_new_tokens, use_cache=True)
            answer = output_ids[:, input_token_len:]
            print(f"{len(images)} image(s) -> {sum(crops_per_image)} crops, prompt {input_token_len} ids "
                  f"({input_token_len + 168 * sum(crops_per_image)} decoder rows) -> answer ids {answer[0].tolist()}")


if __name__ == "__main__":
    main()
