"""CPU: the plugin classes are real `nn.Module`s over the native weights: parameters share storage with what the
kernels read, state_dict speaks the reference's HF keys, load_state_dict is in place; the Megatron shim's train_step takes the
batch from the reference's own forward_step_func."""
import types

import pytest
import torch
from torch import nn

from leopard_b200 import hf as HF
from leopard_b200.config import leopard_llava, reduced
from leopard_b200.native_weights import load_hf
from leopard_b200.weights import SyntheticCheckpoint, tensor_specs


@pytest.fixture
def cpu_model(monkeypatch):
    cfg = reduced(leopard_llava(), vit_layers=1, dec_layers=1, vocab=320)
    ck = SyntheticCheckpoint(cfg, seed=4)
    monkeypatch.setattr(HF, "PrefillEngine", lambda cfg_, native: types.SimpleNamespace(cfg=cfg_, w=native, device=native.device))
    m = HF.myLlavaForConditionalGeneration(cfg)
    m._bind(load_hf(cfg, lambda n: ck[n], "cpu"))
    return cfg, ck, m


def test_llava_shim_is_an_nn_module_over_the_native_weights(cpu_model):
    cfg, ck, m = cpu_model
    assert isinstance(m, nn.Module) and m.eval() is m and m.training is False
    params = dict(m.named_parameters())
    assert params and all(not p.requires_grad for p in params.values())
    # parameters are the tensors the kernels read (same storage), in kernel layout
    assert params["native.dec__0__qkv_w"].data_ptr() == m._native.named["dec.0.qkv_w"].data_ptr()
    n_model = sum(int(torch.tensor(s).prod()) for _, s, _ in tensor_specs(cfg))
    n_native = sum(p.numel() for p in m.parameters())
    assert n_model <= n_native < 1.15 * n_model                     # + head padding (72 -> 80) and the patch K padding
    # state_dict: the reference's HF keys and values
    sd = m.state_dict()
    for k in ("language_model.model.layers.0.self_attn.k_proj.weight", "multi_modal_projector.linear_1.bias",
              "language_model.lm_head.weight", "language_model.model.layers.0.mlp.up_proj.weight"):
        assert torch.equal(sd[k].to(torch.bfloat16).cpu(), ck[k]), k
    # load_state_dict: in place
    other = SyntheticCheckpoint(cfg, seed=9)
    ptr = m._native.named["dec.0.o_w"].data_ptr()
    m.load_state_dict({k: other[k] for k in other.keys()})
    assert m._native.named["dec.0.o_w"].data_ptr() == ptr
    assert torch.equal(m._native.named["dec.0.o_w"], other["language_model.model.layers.0.self_attn.o_proj.weight"])
    assert torch.equal(m.state_dict()["language_model.model.norm.weight"].to(torch.bfloat16), other["language_model.model.norm.weight"])
    with pytest.raises(Exception):
        m.to("cpu")
    assert m.to(torch.float16) is m and m.half() is m


def test_train_step_takes_its_batch_from_the_scripts_forward_step():
    """Megatron's train_step(forward_step_func, data_iterator, model, optimizer, ...): the reference's forward_step
    (examples/llava/pretrain_megatron_llava.py:124-160) is called as is; the native trainer receives exactly its batch."""
    from functools import partial

    from leopard_b200 import megatron as M

    def get_batch(it):                                               # the script's get_batch, reduced to its outputs
        data = next(it)
        tokens, labels, weights = data["input_ids"].long(), data["labels"].long(), data["weights"].float()
        return tokens, labels, weights, None, torch.arange(tokens.shape[1])[None], data["image"]

    def forward_step(data_iterator, model):                           # :124-160
        tokens, labels, loss_mask, attention_mask, position_ids, images = get_batch(data_iterator)
        output_tensor = model(tokens, position_ids, attention_mask, images=images)
        return output_tensor, partial(M.loss_func, loss_mask, labels)

    seen = {}

    class Trainer:
        skipped_last = False

        def step(self, tokens, labels, loss_mask, images):
            seen.update(tokens=tokens, labels=labels, loss_mask=loss_mask, images=images)
            return 1.5, 0.25

    batch = M.mm_collate_fn([{"input_ids": torch.tensor([5, -200, -200, 7, 9]), "labels": torch.tensor([-100, -100, -100, 7, 9]),
                              "weights": torch.tensor([0., 0., 0., 1., 1.]), "preferences": torch.zeros(5),
                              "image": torch.zeros(1, 3, 364, 364, dtype=torch.bfloat16)}])
    model = [types.SimpleNamespace(engine=None)]
    loss, skipped, norm, zeros = M.train_step(forward_step, iter([batch]), model, Trainer())
    assert loss == {"lm loss": 1.5} and skipped == 0 and norm == 0.25 and zeros is None
    assert torch.equal(seen["tokens"], batch["input_ids"]) and torch.equal(seen["labels"], batch["labels"])
    assert torch.equal(seen["loss_mask"], batch["weights"]) and seen["images"] is batch["image"]
    with pytest.raises(M._lib.NativeError):                          # a forward_step that never calls the model is rejected loudly
        M.train_step(lambda it, m: (None, None), iter([batch]), model, Trainer())
