"""The SFT step's host sequencing (leopard_b200/train.py) executed on the CPU with every native op replaced by its
contract restatement (oracle/native_ops.py), against autograd over the model oracle (oracle/train.py).

This pins everything train.py decides — which operand is transposed, which column slice is which head, the interleaved
gate/up layout, the loss shift and weights, where each gradient lands in the flat buffer, bucket order, the optimizer
call — independently of the CUDA kernels, whose own parity is checked per kernel on the GPU (tools/gpu_check_train.py)."""
import math
import types

import pytest
import torch

from leopard_b200 import train as T
from leopard_b200.config import leopard_llava, reduced
from leopard_b200.native_weights import load_hf
from leopard_b200.weights import SyntheticCheckpoint
from oracle import model as OM
from oracle import native_ops
from oracle import train as OT


def _narrow(cfg):
    """Same architecture at a quarter of the widths (decoder 1024 / 8 q heads x 128 / 2 kv heads / ffn 2048, tower 4 heads x 72):
    train.py's sequencing does not depend on the widths, the kernels' width-specific paths are covered under the emulator
    (tests/test_train_kernels_emulated.py), and the CPU suite stays short."""
    from dataclasses import replace
    return replace(cfg, dec=replace(cfg.dec, hidden=1024, heads=8, kv_heads=2, ffn=2048),
                   vit=replace(cfg.vit, hidden=288, heads=4, ffn=512))


class _FakeEngine:
    """The two PrefillEngine methods the trainer calls, restated on the CPU."""

    def __init__(self, cfg, ck, nw):
        self.cfg, self.ck, self.w, self.device = cfg, ck, nw, torch.device("cpu")

    def vit(self, images, shuffle=True):
        with torch.no_grad():
            out = OM.pixel_shuffle(OM.vit_forward(self.ck, self.cfg, images.float()), self.cfg.shuffle)
        return out.reshape(-1, out.shape[-1]).to(torch.bfloat16)

    def merge(self, ids, feats, image_id, rows_per_slot, out=None):
        emb = self.w.embed[ids.clamp(min=0)].clone()
        slots = ids == image_id
        if feats is not None:
            emb[slots] = feats
        out.copy_(emb)
        return out, int(slots.sum())


@pytest.fixture(params=["narrow"])
def cpu_trainer(request, monkeypatch):
    cfg = reduced(leopard_llava(), vit_layers=1, dec_layers=2, vocab=324)     # 324 -> padded to 328: padding rows exist
    if request.param == "narrow":
        cfg = _narrow(cfg)
    ck = SyntheticCheckpoint(cfg, seed=11)
    nw = load_hf(cfg, lambda n: ck[n], device="cpu")
    monkeypatch.setattr(T, "ops", native_ops)
    monkeypatch.setattr(T._lib, "lib", lambda: types.SimpleNamespace(lp_grad_norm_workspace_bytes=lambda: 4096))
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    return cfg, ck, nw, _FakeEngine(cfg, ck, nw)


def _batch(cfg, b, n_text, seed):
    g = torch.Generator().manual_seed(seed)
    rows = cfg.tokens_per_crop
    toks, labs, wts = [], [], []
    for _ in range(b):
        text = torch.randint(0, cfg.dec.vocab, (n_text,), generator=g)
        t = torch.cat([text[:4], torch.full((rows,), cfg.megatron_image_index), text[4:]])
        lab = t.clone()
        lab[t < 0] = -100
        w = torch.zeros(t.shape)
        w[-6:] = 1.0
        toks.append(t), labs.append(lab), wts.append(w)
    images = torch.rand((b, 3, 364, 364), generator=g) * 2 - 1
    return torch.stack(toks), torch.stack(labs), torch.stack(wts), images


def _native_layout(cfg, grads, name):
    """Oracle gradients (HF names) re-laid out like NativeWeights.named[name]."""
    d = cfg.dec
    LM, P = OM.LM, OM.P
    if name == "embed":
        return grads[LM + "embed_tokens.weight"]
    if name == "dec.final_g":
        return grads[LM + "norm.weight"]
    if name == "dec.lm_head":
        g = grads["language_model.lm_head.weight"]
        return torch.nn.functional.pad(g, (0, 0, 0, (d.vocab + 7) // 8 * 8 - d.vocab))
    if name.startswith("proj."):
        return grads[P + {"w1": "linear_1.weight", "b1": "linear_1.bias", "w2": "linear_2.weight", "b2": "linear_2.bias"}[name[5:]]]
    _, i, kind = name.split(".")
    b = f"{LM}layers.{i}."
    if kind == "qkv_w":
        return torch.cat([grads[b + f"self_attn.{n}_proj.weight"] for n in "qkv"], 0)
    if kind == "gateup_w":
        g = grads[b + "mlp.gate_proj.weight"].reshape(d.ffn // 32, 1, 32, d.hidden)
        u = grads[b + "mlp.up_proj.weight"].reshape(d.ffn // 32, 1, 32, d.hidden)
        return torch.cat([g, u], 1).reshape(2 * d.ffn, d.hidden)
    return grads[b + {"o_w": "self_attn.o_proj.weight", "down_w": "mlp.down_proj.weight", "ln1_g": "input_layernorm.weight",
                      "ln2_g": "post_attention_layernorm.weight"}[kind]]


@pytest.mark.parametrize("cpu_trainer", ["narrow", "full"], indirect=True)     # once at Leopard-LLaVA's real widths as well
def test_sft_step_sequencing_matches_autograd_oracle(cpu_trainer):
    cfg, ck, nw, eng = cpu_trainer
    tokens, labels, weights, images = _batch(cfg, b=2, n_text=15, seed=3)        # s = 169 + 15 = 184 (multiple of 8)
    ock = OT.TrainableCheckpoint(ck, trainable=lambda n: not n.startswith("vision_tower."))
    ref_loss, ref = OT.sft_loss_and_grads(ock, cfg, tokens, labels, weights, images)
    ref_norm = OT.global_grad_norm(ref.values())

    before = {n: t.clone() for n, t in nw.named.items()}
    tr = T.SFTTrainer(eng, lr=1e-5, clip_grad=1.0)
    assert len(tr.buckets) == 1 + cfg.dec.layers + 2 and tr.buckets[-1][1] == tr.flat_grad.numel()
    loss, norm = tr.step(tokens, labels, weights, images)

    assert abs(loss - ref_loss) < 0.02 * abs(ref_loss), (loss, ref_loss)
    assert abs(norm - ref_norm) < 0.03 * ref_norm, (norm, ref_norm)
    worst = {}
    for n in tr.names:
        want, got = _native_layout(cfg, ref, n).float(), tr.grad(n).float()
        assert want.shape == got.shape, n
        worst[n] = float((got - want).norm() / (want.norm() + 1e-12))
    bad = {n: e for n, e in worst.items() if e > 0.05}
    assert not bad, bad
    # vocabulary padding rows and unused embedding rows stay exactly zero, and so do their updates
    assert float(tr.grad("dec.lm_head")[cfg.dec.vocab:].abs().max()) == 0.0
    assert torch.equal(nw.named["dec.lm_head"][cfg.dec.vocab:], before["dec.lm_head"][cfg.dec.vocab:])
    # the optimizer moved every trainable tensor, by about lr per element (first Adam step), in the descent direction
    coeff = OT.clip_coefficient(ref_norm, 1.0)
    for n in ("dec.1.o_w", "dec.0.ln1_g", "proj.b1", "dec.final_g"):
        delta = nw.named[n].float() - before[n].float()
        if nw.named[n].dtype == torch.float32:
            g = _native_layout(cfg, ref, n) * coeff
            big = g.abs() > 0.05 * g.abs().max()            # away from zero the bf16 noise cannot flip the sign
            assert torch.allclose(delta[big], -1e-5 * torch.sign(g[big]), rtol=0.05, atol=0), n   # m / (sqrt(v) + eps)
        else:
            assert float(delta.abs().max()) > 0.0, n
    # frozen tower: untouched
    assert tr.iteration == 1 and not any(n.startswith("vit") for n in tr.names)


def test_flash_backward_call_is_equivalent_to_the_materialised_path(cpu_trainer):
    """train.py's two attention-backward routes (per-head materialised tiles vs one lp_attention_bwd call) must be fed
    consistently: with the op contracts in place of the kernels they give the same gradients."""
    cfg, eng = _small_cpu_engine()
    tokens, labels, weights = _text_batch(cfg, seed=4)
    grads = {}
    for mode in ("materialized", "flash"):
        tr = T.SFTTrainer(eng, lr=0.0, attn_bwd=mode)          # lr 0: the second run sees the same weights
        tr.step(tokens, labels, weights, None)
        grads[mode] = tr.flat_grad.clone()
    rel = float((grads["flash"] - grads["materialized"]).norm() / grads["materialized"].norm())
    assert rel < 0.02, rel


def test_projector_only_stage_skips_decoder_gradients(cpu_trainer):
    cfg, ck, nw, eng = cpu_trainer
    tokens, labels, weights, images = _batch(cfg, b=1, n_text=15, seed=5)
    ock = OT.TrainableCheckpoint(ck, trainable=lambda n: n.startswith("multi_modal_projector."))
    ref_loss, ref = OT.sft_loss_and_grads(ock, cfg, tokens, labels, weights, images)
    before = {n: t.clone() for n, t in nw.named.items()}
    tr = T.SFTTrainer(eng, lr=1e-5, freeze_llm=True)
    assert tr.names == ["proj.w2", "proj.b2", "proj.w1", "proj.b1"] and len(tr.buckets) == 1
    loss, _ = tr.step(tokens, labels, weights, images)
    assert abs(loss - ref_loss) < 0.02 * abs(ref_loss)
    for n in tr.names:
        want, got = _native_layout(cfg, ref, n).float(), tr.grad(n).float()
        assert float((got - want).norm() / want.norm()) < 0.05, n
    for n, t in nw.named.items():
        if not n.startswith("proj."):
            assert torch.equal(t, before[n]), n


def test_shifted_targets_and_schedule_helpers():
    labels = torch.tensor([[5, 6, 7, 8], [1, 2, 3, 4]])
    mask = torch.tensor([[0.0, 0.0, 1.0, 1.0], [1.0, 0.0, 0.0, 1.0]])
    tgt, w = T.shifted_targets(labels, mask, dp_world=2)
    assert tgt.tolist() == [6, 7, 8, -1, 2, 3, 4, -1]
    assert torch.allclose(w, torch.tensor([0, 1, 1, 0, 0, 0, 1, 0]) / 3 / 2)
    # same loss as the oracle's loss_func restatement for random logits
    g = torch.Generator().manual_seed(0)
    logits = torch.randn((2, 4, 11), generator=g)
    ref, dref = OT.masked_lm_loss(logits, labels, mask)
    tgt, w = T.shifted_targets(labels, mask)
    rows, d = native_ops.ce_fwd_bwd(logits.reshape(8, 11), tgt, w, 11)
    assert abs(float(torch.dot(rows, w)) - float(ref)) < 1e-6
    assert torch.allclose(d.float().view(2, 4, 11), dref, atol=4e-3)
    assert T.cosine_lr(5, 1e-3, 1e-4, 10, 100) == OT.cosine_lr(5, 1e-3, 1e-4, 10, 100)
    assert T.cosine_lr(60, 1e-3, 1e-4, 10, 100) == OT.cosine_lr(60, 1e-3, 1e-4, 10, 100)


# ---- data parallel: two gloo ranks, one micro-batch each, gradients averaged bucket by bucket ----------------------------
def _text_batch(cfg, seed):
    g = torch.Generator().manual_seed(seed)
    tokens = torch.randint(0, cfg.dec.vocab, (1, 16), generator=g)
    weights = torch.zeros(1, 16)
    weights[0, 8 + seed % 3:] = 1.0            # ranks weight a different number of tokens
    return tokens, tokens.clone(), weights


def _fingerprint(tr, flat=None):
    """Three numbers per trainable tensor (sum, sum of squares, projection on a fixed vector): small enough for a queue."""
    flat = tr.flat_grad if flat is None else flat
    out = []
    for n in tr.names:
        off, cnt = tr.layout[n]
        g = flat[off:off + cnt].double()
        out.append([float(g.sum()), float(g.pow(2).sum()), float(torch.dot(g, torch.cos(torch.arange(cnt, dtype=torch.float64) * 0.37)))])
    return torch.tensor(out, dtype=torch.float64)


def _patch_cpu():
    T.ops = native_ops
    T._lib.lib = lambda: types.SimpleNamespace(lp_grad_norm_workspace_bytes=lambda: 4096)
    torch.cuda.is_available = lambda: True


_SMALL_CACHE = {}


def _small_cpu_engine():
    """Fresh native weights every call (trainers update them in place); the synthetic tensors are drawn once per process."""
    cfg = _narrow(reduced(leopard_llava(), vit_layers=1, dec_layers=1, vocab=320))
    ck = SyntheticCheckpoint(cfg, seed=5)

    def get(name):
        if name not in _SMALL_CACHE:
            _SMALL_CACHE[name] = ck[name]
        return _SMALL_CACHE[name]

    nw = load_hf(cfg, get, device="cpu")
    return cfg, _FakeEngine(cfg, ck, nw)


def _dp_train_worker(rank, world, port, q):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _patch_cpu()
        cfg, eng = _small_cpu_engine()
        tr = T.SFTTrainer(eng, lr=1e-5, clip_grad=1.0)
        assert tr.world == world
        loss, norm = tr.step(*_text_batch(cfg, seed=rank), None)
        q.put((rank, loss, norm, _fingerprint(tr), eng.w.named["dec.0.o_w"].float().sum().item()))
    finally:
        dist.destroy_process_group()


def test_data_parallel_step_averages_gradients_world2_gloo(cpu_trainer):
    import os
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=600) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (_, l0, n0, g0, w0), (_, l1, n1, g1, w1) = got
    assert torch.equal(g0, g1) and n0 == n1 and w0 == w1          # ranks hold the same reduced gradient and stay in sync
    assert l0 == l1                  # the logged loss is the data-parallel average (pretrain_megatron_llava.py:119)
    # single-rank gradients of the two micro-batches, averaged
    singles, single_losses = [], []
    for seed in (0, 1):
        cfg, eng = _small_cpu_engine()
        tr = T.SFTTrainer(eng, lr=1e-5)
        loss, _ = tr.step(*_text_batch(cfg, seed=seed), None)
        single_losses.append(loss)
        singles.append(tr.flat_grad.clone())
    assert single_losses[0] != single_losses[1] and abs(0.5 * sum(single_losses) - l0) < 1e-3 * abs(l0)
    want = _fingerprint(tr, 0.5 * (singles[0] + singles[1]))
    # sums and projections are linear in the gradient: they must match the average up to the bf16 noise of the
    # pre-scaled loss weights; the square sums agree to the same order
    scale = want[:, 1].sqrt()[:, None] + 1e-30
    assert float(((g0[:, [0, 2]] - want[:, [0, 2]]) / scale).abs().max()) < 0.05
    assert float(((g0[:, 1] - want[:, 1]) / (want[:, 1] + 1e-30)).abs().max()) < 0.05


def test_optimizer_state_round_trip_resumes_identically(cpu_trainer):
    """Two steps in one trainer == one step, state_dict -> fresh model + trainer -> load_state_dict, second step."""
    cfg, eng = _small_cpu_engine()
    batch = _text_batch(cfg, seed=2)
    tr = T.SFTTrainer(eng, lr=1e-4)
    tr.step(*batch, None)
    sd = tr.state_dict()
    assert sd["iteration"] == 1 and set(sd["main"]) == set(tr.names)
    loss_a, norm_a = tr.step(*batch, None)

    cfg2, eng2 = _small_cpu_engine()                     # fresh (untrained) weights; everything comes from the state
    tr2 = T.SFTTrainer(eng2, lr=1e-4)
    tr2.load_state_dict(sd)
    loss_b, norm_b = tr2.step(*batch, None)
    assert (loss_a, norm_a) == (loss_b, norm_b) and tr2.iteration == 2
    for n in tr.names:
        assert torch.equal(eng.w.named[n], eng2.w.named[n]), n
    assert torch.equal(tr.flat_master, tr2.flat_master) and torch.equal(tr.flat_v, tr2.flat_v)


def test_gradient_accumulation_over_micro_batches(cpu_trainer):
    """step([mb0, mb1]) == the average of the two single-micro-batch gradients; reported loss = mean of the two."""
    cfg, eng = _small_cpu_engine()
    mbs = [_text_batch(cfg, seed=s) for s in (0, 1)]
    singles, losses = [], []
    for mb in mbs:
        tr = T.SFTTrainer(eng, lr=0.0)
        loss, _ = tr.step(*mb, None)
        singles.append(tr.flat_grad.clone())
        losses.append(loss)
    tr = T.SFTTrainer(eng, lr=0.0)
    loss, norm = tr.step([m[0] for m in mbs], [m[1] for m in mbs], [m[2] for m in mbs], [None, None])
    want = 0.5 * (singles[0] + singles[1])
    assert abs(loss - 0.5 * (losses[0] + losses[1])) < 1e-3 * abs(loss)
    assert float((tr.flat_grad - want).norm() / want.norm()) < 0.01
    assert abs(norm - float(want.double().norm())) < 1e-4 * norm and tr.iteration == 1     # fp64: 2e8 addends


def test_unfrozen_vision_tower_backward_matches_autograd_oracle(cpu_trainer):
    """freeze_vision_tower=False: the tower's forward is re-sequenced in train.py (saved layer inputs) and its backward runs
    through LayerNorm / GELU / bias / position-embedding / patch-embedding gradients and the per-(crop, head) attention
    backward on the zero-padded 80-wide head layout."""
    cfg = _narrow(reduced(leopard_llava(), vit_layers=2, dec_layers=1, vocab=320))
    ck = SyntheticCheckpoint(cfg, seed=13)
    nw = load_hf(cfg, lambda n: ck[n], device="cpu")
    eng = _FakeEngine(cfg, ck, nw)
    tokens, labels, weights, images = _batch(cfg, b=1, n_text=15, seed=6)
    ock = OT.TrainableCheckpoint(ck)
    ref_loss, ref = OT.sft_loss_and_grads(ock, cfg, tokens, labels, weights, images)
    ref_hf = {k: v for k, v in ref.items()}

    tr = T.SFTTrainer(eng, lr=0.0, freeze_vision_tower=False)
    assert len(tr.buckets) == 1 + cfg.dec.layers + 3
    loss, norm = tr.step(tokens, labels, weights, images)
    assert abs(loss - ref_loss) < 0.02 * abs(ref_loss)
    assert abs(norm - OT.global_grad_norm(ref.values())) < 0.03 * norm

    # bring the trainer's gradients back to HF names with the same inverse re-layout the checkpoint export uses
    from leopard_b200.native_weights import named_to_hf
    got = named_to_hf(cfg, {n: tr.grad(n) for n in tr.names})
    bad = {}
    for name, want in ref_hf.items():
        g = got[name].float()
        assert g.shape == want.shape, name
        if name.endswith("self_attn.k_proj.bias"):
            # softmax is invariant to a per-query constant, so this gradient is exactly 0 in exact arithmetic: what is left
            # on both sides is rounding noise — it only has to be small next to the query bias's gradient
            qn = float(ref_hf[name.replace("k_proj", "q_proj")].norm())
            assert float(want.norm()) < 1e-3 * qn and float(g.norm()) < 0.05 * qn, (name, float(g.norm()), qn)
            continue
        e = float((g - want).norm() / (want.norm() + 1e-12))
        if e > 0.06:
            bad[name] = e
    assert not bad, bad
    # the zero padding of the 80-wide head layout and of the patch-embedding K dimension receives exactly zero gradient
    qg = tr.grad("vit.1.qkv_w").view(3, cfg.vit.heads, 80, -1)
    assert float(qg[:, :, 72:].abs().max()) == 0.0
    assert float(tr.grad("vit.0.qkv_b").view(3, cfg.vit.heads, 80)[:, :, 72:].abs().max()) == 0.0
    assert float(tr.grad("vit.patch_w")[:, cfg.vit.patch_k:].abs().max()) == 0.0
    # the flash route (one lp_attention_bwd call per layer, (72, 80) head layout) is fed consistently with the per-head one
    tr2 = T.SFTTrainer(eng, lr=0.0, freeze_vision_tower=False, attn_bwd="flash")
    tr2.step(tokens, labels, weights, images)
    assert float((tr2.flat_grad - tr.flat_grad).norm() / tr.flat_grad.norm()) < 0.02


def test_non_finite_gradients_skip_the_iteration(cpu_trainer):
    """inf / NaN in the gradients: weights, moments and the step counter stay untouched and train_step reports the skip
    (Megatron's found-inf path)."""
    from leopard_b200 import megatron
    cfg, eng = _small_cpu_engine()
    tokens, labels, weights = _text_batch(cfg, seed=3)
    tr = T.SFTTrainer(eng, lr=1e-4)
    before = {n: t.clone() for n, t in eng.w.named.items()}
    eng.w.named["dec.lm_head"][5, 7] = float("inf")            # poisons the logits of every row
    poisoned = eng.w.named["dec.lm_head"].clone()
    batch = {"input_ids": tokens, "labels": labels, "weights": weights, "image": torch.zeros(1, 10, dtype=torch.bfloat16)}
    loss, skipped, norm, _ = megatron.train_step(None, iter([batch]), [object()], tr)
    assert skipped == 1 and not math.isfinite(norm) and tr.iteration == 0
    assert torch.equal(eng.w.named["dec.lm_head"], poisoned) and float(tr.flat_m.abs().max()) == 0.0
    for n in ("dec.0.o_w", "proj.b1", "embed"):
        assert torch.equal(eng.w.named[n], before[n]), n
    eng.w.named["dec.lm_head"].copy_(before["dec.lm_head"])     # healthy again: the next step counts as iteration 1
    loss, skipped, norm, _ = megatron.train_step(None, iter([batch]), [object()], tr)
    assert skipped == 0 and math.isfinite(norm) and tr.iteration == 1


def test_no_recompute_mode_gives_the_same_gradients(cpu_trainer):
    """recompute='none' (activations kept, no second forward per layer) against recompute='full', tower trained as well."""
    cfg, ck, nw, eng = cpu_trainer
    tokens, labels, weights, images = _batch(cfg, b=2, n_text=15, seed=8)
    grads, losses = {}, {}
    for mode in ("full", "none"):
        tr = T.SFTTrainer(eng, lr=0.0, freeze_vision_tower=False, recompute=mode)
        losses[mode], _ = tr.step(tokens, labels, weights, images)
        grads[mode] = tr.flat_grad.clone()
    # the two differ only by where SwiGLU is rounded in the forward that produces the loss (fused epilogue vs materialised)
    assert abs(losses["full"] - losses["none"]) < 5e-3 * abs(losses["full"])
    assert float((grads["none"] - grads["full"]).norm() / grads["full"].norm()) < 0.02


def test_checkpoint_loaded_after_the_trainer_was_built_is_what_gets_trained(cpu_trainer, monkeypatch):
    """Megatron's order is setup_model_and_optimizer, then load_checkpoint: GPTModel.load_state_dict must load IN PLACE so
    that the trainer's tensors, its fp32 main copies and the exported checkpoint all refer to the loaded weights."""
    from leopard_b200 import megatron as M
    from leopard_b200.native_weights import hf_to_megatron
    cfg, ck, nw, eng = cpu_trainer
    model = M.GPTModel.__new__(M.GPTModel)
    model.cfg, model.engine, model.input_tensor, model._hf_getter = cfg, eng, None, (lambda n: ck[n])
    tr = T.SFTTrainer(eng, lr=1e-2, clip_grad=1.0)                       # built BEFORE the checkpoint is loaded
    other = SyntheticCheckpoint(cfg, seed=99)
    state = {"language_model": M.nest_language_model_state(hf_to_megatron(cfg, lambda n: other[n]))}
    ptr = eng.w.named["dec.0.o_w"].data_ptr()
    model.load_state_dict(state)
    assert model.engine is eng and eng.w.named["dec.0.o_w"].data_ptr() == ptr            # same objects, new values
    want = load_hf(cfg, lambda n: other[n], device="cpu").named
    assert all(torch.equal(eng.w.named[n], want[n]) for n in want)
    assert torch.equal(tr._view(tr.flat_master, tr.master_layout, "dec.0.o_w"), want["dec.0.o_w"].float())
    loaded = model.state_dict_for_save_checkpoint()["language_model"]["encoder"]["layers.0.self_attention.dense.weight"].clone()
    tokens, labels, weights, images = _batch(cfg, b=1, n_text=15, seed=4)
    tr.step(tokens, labels, weights, images)
    trained = model.state_dict_for_save_checkpoint()["language_model"]["encoder"]["layers.0.self_attention.dense.weight"]
    assert not torch.equal(trained, loaded)                                # the export sees the update
    assert torch.equal(trained, eng.w.named["dec.0.o_w"].cpu())            # ... of the tensors forward() uses
