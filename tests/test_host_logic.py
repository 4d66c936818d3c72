"""not gpu: host-side logic — weight-layout conversions, native re-layout (pure data movement), DP sharding over a
world_size-2 gloo group, bench workload bookkeeping."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from leopard_b200 import dp
from leopard_b200.config import leopard_llava, reduced
from leopard_b200.native_weights import HEAD_STRIDE, PATCH_K_PAD, hf_to_megatron, load_hf, megatron_to_hf
from leopard_b200.weights import SyntheticCheckpoint, tensor_specs


def tiny_cfg():
    return reduced(leopard_llava(), vit_layers=1, dec_layers=1, vocab=512)


def test_megatron_hf_roundtrip_and_group_interleave():
    cfg = tiny_cfg()
    ck = SyntheticCheckpoint(cfg)
    meg = hf_to_megatron(cfg, lambda n: ck[n])
    back = megatron_to_hf(cfg, meg)
    for name, _, _ in tensor_specs(cfg):
        assert torch.equal(back[name], ck[name]), name
    # fused QKV rows are grouped per kv head: [4 q heads | k | v] (transformer.py:713-738)
    qkv = meg["encoder.layers.0.self_attention.query_key_value.weight"]
    assert qkv.shape == ((32 + 16) * 128, 4096)
    q = ck["language_model.model.layers.0.self_attn.q_proj.weight"]
    k = ck["language_model.model.layers.0.self_attn.k_proj.weight"]
    g = 3
    assert torch.equal(qkv[g * 768: g * 768 + 512], q[g * 512:(g + 1) * 512])
    assert torch.equal(qkv[g * 768 + 512: g * 768 + 640], k[g * 128:(g + 1) * 128])
    h4 = meg["encoder.layers.0.mlp.dense_h_to_4h.weight"]
    assert torch.equal(h4[:14336], ck["language_model.model.layers.0.mlp.gate_proj.weight"])
    # ViT qkv interleave [head, (q,k,v), 72]
    vq = meg["vision_tower.transformer.layers.0.self_attention.linear_qkv.weight"].view(16, 3, 72, 1152)
    assert torch.equal(vq[5, 1], ck["vision_tower.vision_model.encoder.layers.0.self_attn.k_proj.weight"][5 * 72:6 * 72])


def test_native_layout_is_pure_relayout():
    cfg = tiny_cfg()
    ck = SyntheticCheckpoint(cfg)
    nw = load_hf(cfg, lambda n: ck[n], "cpu")
    tensors = {t.data_ptr(): t for t in nw._keep}
    L = nw.vit_layers[0]
    qkv = tensors[L.qkv_w]
    assert qkv.shape == (3 * 16 * HEAD_STRIDE, 1152)
    k = ck["vision_tower.vision_model.encoder.layers.0.self_attn.k_proj.weight"]
    blk = qkv[16 * HEAD_STRIDE:2 * 16 * HEAD_STRIDE].view(16, HEAD_STRIDE, 1152)
    assert torch.equal(blk[:, :72].reshape(1152, 1152), k) and bool((blk[:, 72:] == 0).all())
    qb = tensors[L.qkv_b].view(3, 16, HEAD_STRIDE)
    assert bool((qb[:, :, 72:] == 0).all())
    pw = tensors[nw.vit.patch_w]
    assert pw.shape == (1152, PATCH_K_PAD) and bool((pw[:, 588:] == 0).all())
    D = nw.dec_layers[0]
    gu = tensors[D.gateup_w].view(14336 // 32, 2, 32, 4096)
    assert torch.equal(gu[:, 0].reshape(14336, 4096), ck["language_model.model.layers.0.mlp.gate_proj.weight"])
    assert torch.equal(gu[:, 1].reshape(14336, 4096), ck["language_model.model.layers.0.mlp.up_proj.weight"])
    qkvd = tensors[D.qkv_w]
    assert qkvd.shape == (6144, 4096)
    assert torch.equal(qkvd[4096:5120], ck["language_model.model.layers.0.self_attn.k_proj.weight"])
    assert nw.vocab_padded == 512


def test_vocab_padding_for_mistral():
    from dataclasses import replace
    from leopard_b200.config import leopard_idefics2_decoder
    cfg = tiny_cfg()
    cfg = replace(cfg, dec=replace(leopard_idefics2_decoder(), layers=1, vocab=1003), image_token_index=1001)
    ck = SyntheticCheckpoint(cfg)
    nw = load_hf(cfg, lambda n: ck[n], "cpu")
    assert nw.vocab_padded == 1008


def test_assignment_is_balanced_and_deterministic():
    costs = [dp.request_cost(n, 64 + 169 * n) for n in [50, 1, 1, 28, 17, 3, 3, 40, 9, 9, 2, 30]]
    bins = dp.assign(costs, 4)
    assert sorted(i for b in bins for i in b) == list(range(12))
    loads = [sum(costs[i] for i in b) for b in bins]
    assert max(loads) / (sum(loads) / 4) < 1.15
    naive = [sum(costs[i] for i in b) for b in dp.contiguous(12, 4)]
    assert max(loads) < max(naive)
    assert dp.assign(costs, 4) == bins
    assert dp.contiguous(10, 4) == [[0, 1, 2], [3, 4, 5], [6, 7, 8], [9]]


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        reqs = None
        n_req = 5
        if rank == 0:
            reqs = []
            for k, n in enumerate([3, 0, 1, 2, 1]):
                ids = torch.randint(0, 1000, (10 + k,), generator=g)
                pix = None if n == 0 else torch.randn(n, 3, 4, 4, generator=g)
                reqs.append((ids, pix))
        local, mine = dp.scatter_requests(reqs, "cpu", src=0)
        # stand-in for the per-request result: [sum(ids), n_crops, pixel checksum]
        res = torch.tensor([[float(i.sum()), 0.0 if p is None else float(p.shape[0]),
                             0.0 if p is None else float(p.sum())] for i, p in local]).reshape(len(local), 3)
        allres = dp.gather_results(res, mine, n_req)
        q.put((rank, mine, allres))
    finally:
        dist.destroy_process_group()


def test_scatter_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort(key=lambda t: t[0])
    (r0, m0, a0), (r1, m1, a1) = got
    assert sorted(m0 + m1) == [0, 1, 2, 3, 4] and m0 and m1
    assert torch.equal(a0, a1)
    g = torch.Generator().manual_seed(0)
    for k, n in enumerate([3, 0, 1, 2, 1]):
        ids = torch.randint(0, 1000, (10 + k,), generator=g)
        pix = None if n == 0 else torch.randn(n, 3, 4, 4, generator=g)
        assert float(a0[k, 0]) == float(ids.sum()) and float(a0[k, 1]) == float(n)
        if pix is not None:
            assert abs(float(a0[k, 2]) - float(pix.sum())) < 1e-4


def test_bench_workload_bookkeeping():
    import bench
    cfg = leopard_llava()
    pix, ids, n_crops, S, per_image = bench.make_workload(cfg)
    assert per_image == [7, 7, 7, 7] and n_crops == 28 and S == 4796 and ids.shape == (1, 92)
    assert pix.shape == (28, 3, 364, 364) and float(pix.min()) >= -1 and float(pix.max()) <= 1
    fm = bench.flops_model(n_crops, S, cfg)
    assert abs(fm["total"] / 1e12 - 90.5) < 0.3  # 90.5 TFLOP per cfg3 request


def test_from_pretrained_reads_hf_checkpoint_dir(tmp_path):
    """config.json + safetensors in the HF layout the reference converter writes -> getter -> native re-layout (CPU)."""
    import json
    from safetensors.torch import save_file
    from leopard_b200 import hf
    cfg = tiny_cfg()
    ck = SyntheticCheckpoint(cfg)
    tensors = {n: ck[n].contiguous() for n, _, _ in tensor_specs(cfg)}
    names = sorted(tensors)
    save_file({n: tensors[n] for n in names[: len(names) // 2]}, str(tmp_path / "model-00001-of-00002.safetensors"))
    save_file({n: tensors[n] for n in names[len(names) // 2:]}, str(tmp_path / "model-00002-of-00002.safetensors"))
    (tmp_path / "config.json").write_text(json.dumps({
        "image_token_index": cfg.image_token_index, "projector_hidden_act": "gelu",
        "text_config": {"hidden_size": 4096, "intermediate_size": 14336, "num_attention_heads": 32,
                        "num_key_value_heads": 8, "num_hidden_layers": 1, "vocab_size": 512, "rope_theta": 500000.0,
                        "rope_scaling": {"factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                                         "original_max_position_embeddings": 8192, "rope_type": "llama3"},
                        "rms_norm_eps": 1e-5},
        "vision_config": {"hidden_size": 1152, "intermediate_size": 4304, "num_attention_heads": 16,
                          "num_hidden_layers": 1, "image_size": 364, "patch_size": 14, "layer_norm_eps": 1e-6}}))
    m = hf.myLlavaForConditionalGeneration.from_pretrained(str(tmp_path), torch_dtype=torch.float32)
    assert m.config.dec.layers == 1 and m.config.vit.layers == 1 and m.config.dec.rope_scaling.factor == 8.0
    nw = load_hf(m.config, m._pending, "cpu")
    assert torch.equal(nw.embed, ck["language_model.model.embed_tokens.weight"])
    from leopard_b200._lib import NativeError
    with pytest.raises(NativeError, match="CUDA"):
        m.to("cpu")


def test_native_kv_cache_is_the_legacy_tuple_the_reference_indexes():
    """`past_key_values[l][k|v]` must look like transformers' legacy cache — [1, kv_heads, S, head_dim] per layer, as
    read at evaluations/models/llava_multiimg_siglip_anyres.py:294 — while staying a VIEW of the native buffer."""
    import torch
    from leopard_b200.hf import NativeKVCache
    layers, cap, kvh, hd, S = 3, 16, 8, 128, 11
    kv = torch.arange(layers * 2 * cap * kvh * hd, dtype=torch.float32).view(layers, 2, cap, kvh * hd)
    pkv = NativeKVCache(kv, S, kvh, hd)
    assert isinstance(pkv, tuple) and len(pkv) == layers and pkv.length == S and pkv.get_seq_length() == S
    k0, v0 = pkv[0]
    assert k0.shape == (1, kvh, S, hd) and v0.shape == (1, kvh, S, hd)
    assert pkv[0][0][:, :, :, 0].shape == (1, kvh, S)            # the reference's mask probe
    # element (head h, row s, dim d) of layer l is kv[l, 0, s, h*hd + d]; and it is a view, not a copy
    assert float(pkv[2][1][0, 5, 7, 9]) == float(kv[2, 1, 7, 5 * hd + 9])
    kv[1, 0, 3, 2 * hd + 4] = -1.0
    assert float(pkv[1][0][0, 2, 3, 4]) == -1.0
    grown = NativeKVCache(kv, S + 1, kvh, hd)
    assert grown[0][0].shape[2] == S + 1 and grown.kv.data_ptr() == pkv.kv.data_ptr()


def test_megatron_checkpoint_nesting_round_trip():
    """state_dict_for_save_checkpoint / load_state_dict plumbing of the Megatron shim (gpt_model.py:115-137,
    vlm_model.py:684-791): flat Megatron keys <-> the nested dict saved under state['model']['language_model']."""
    import torch
    from dataclasses import replace
    from leopard_b200.config import leopard_llava, reduced
    from leopard_b200.megatron import flatten_language_model_state, nest_language_model_state
    from leopard_b200.native_weights import hf_to_megatron, megatron_to_hf
    from leopard_b200.weights import SyntheticCheckpoint
    cfg = reduced(leopard_llava(), vit_layers=1, dec_layers=1, vocab=1024)
    ck = SyntheticCheckpoint(cfg, seed=5)
    flat = hf_to_megatron(cfg, lambda n: ck[n])
    nested = nest_language_model_state(flat)
    assert set(nested) == {"embedding", "vision_tower", "mm_projector", "encoder", "output_layer"}
    assert set(nested["embedding"]) == {"word_embeddings"} and "weight" in nested["embedding"]["word_embeddings"]
    assert "conv1.weight" in nested["vision_tower"]                     # no 'vision_tower.' prefix inside (converter :1313)
    assert "encoder.linear_fc2.bias" in nested["mm_projector"]          # converter reads exactly this key (:1091)
    assert "layers.0.self_attention.query_key_value.weight" in nested["encoder"] and "weight" in nested["output_layer"]
    # every spelling load_state_dict accepts flattens back to the same tensors
    nested["vision_tower"]["transformer.layers.0.self_attention.linear_qkv._extra_state"] = torch.zeros(1)   # TE metadata
    for form in ({"language_model": nested}, nested, dict(flat)):
        back = flatten_language_model_state(form)
        assert set(back) == set(flat) and all(torch.equal(back[k], flat[k]) for k in flat)
    # legacy spellings (vlm_model.py:752-760)
    legacy = {k.replace("_norm", "_layernorm").replace(".self_attention.", ".attention.")
              if k.startswith("encoder.") else k: v for k, v in flat.items()}
    assert "encoder.layers.0.input_layernorm.weight" in legacy and "encoder.final_layernorm.weight" in legacy
    assert set(flatten_language_model_state(legacy)) == set(flat)
    hf = megatron_to_hf(cfg, flatten_language_model_state({"language_model": nested}))
    for name in ("language_model.model.layers.0.self_attn.k_proj.weight", "multi_modal_projector.linear_2.bias",
                 "vision_tower.vision_model.encoder.layers.0.self_attn.q_proj.weight"):
        assert torch.equal(hf[name], ck[name])


def test_eval_prompt_builder_matches_reference_executed_goldens():
    """leopard_b200.prompt.build_eval_prompt == the reference's own lines 412-442 (executed by
    tests/golden/gen_prompt_goldens.py), quirks and failure modes included."""
    import json
    from pathlib import Path
    from leopard_b200.prompt import build_eval_prompt, get_instruction, image_placeholder
    cases = json.loads((Path(__file__).resolve().parent / "golden" / "ref_prompts.json").read_text())
    assert len(cases) >= 9
    for c in cases:
        instr = get_instruction(c["setting"], c["ques_type"])
        if "raises" in c:
            with pytest.raises(IndexError):
                build_eval_prompt(c["question"], instr, c["image_num"], c["tiles"])
            assert c["raises"] == "IndexError"
        else:
            assert build_eval_prompt(c["question"], instr, c["image_num"], c["tiles"]) == c["prompt"], c["question"]
    # one placeholder per crop: a 4-image cfg3 request has 4 x (1 + 6) slots = 28 crops
    p = build_eval_prompt(cases[0]["question"], "", 4, [6, 6, 6, 6])
    assert p.count("<|reserved_special_token_195|>") == 28 and image_placeholder(2, 7) in p


def test_train_sample_builder_matches_reference_executed_goldens():
    """leopard_b200.prompt.build_train_sample == the reference's preprocess_llama3 (executed unmodified by
    tests/golden/gen_train_goldens.py with the toy tokenizer): placeholder layout tiles + [global] with tile tags,
    -200 per feature row, loss weights only on assistant bodies, dropped / failing samples alike."""
    import copy
    import json
    import sys
    from pathlib import Path
    G = Path(__file__).resolve().parent / "golden"
    sys.path.insert(0, str(G))
    from toy_tokenizer import ToyTokenizer
    from leopard_b200.prompt import build_train_sample, train_image_placeholder
    tok = ToyTokenizer()
    cases = json.loads((G / "ref_train_samples.json").read_text())
    assert len(cases) >= 6
    for c in cases:
        kw = dict(num_image=c["num_image"], per_image_token=c["per_image_token"],
                  num_patches_per_image=c["num_patches_per_image"])
        if "raises" in c:
            with pytest.raises(Exception) as ei:
                build_train_sample(copy.deepcopy(c["sources"]), tok, **kw)
            assert type(ei.value).__name__ == c["raises"]
            continue
        got = build_train_sample(copy.deepcopy(c["sources"]), tok, **kw)
        if c["result"] is None:
            assert got is None
            continue
        assert set(got) == set(c["result"])
        for k, v in c["result"].items():
            assert got[k].tolist() == v, (k, c["sources"])
    # 2 tiles + global, 3 slots per crop: <20> sss<22> sss<23> sss<21>
    s = "<|reserved_special_token_250|>" * 3
    assert train_image_placeholder(1, 3, 2) == (f"image 1: <|reserved_special_token_20|>{s}<|reserved_special_token_22|>"
                                                f"{s}<|reserved_special_token_23|>{s}<|reserved_special_token_21|>")


def test_loss_func_matches_reference_executed_goldens(monkeypatch):
    """megatron.loss_func: shift / weight / normalise arithmetic against the reference's own loss_func, executed by
    tests/golden/gen_loss_goldens.py.  The per-token cross entropy is the native kernel on a GPU (parity: test_gpu_ops);
    here it is swapped for a torch stand-in so that only the host arithmetic around it is under test."""
    import importlib.util
    import json
    from pathlib import Path

    from leopard_b200 import megatron, ops

    g = Path(__file__).parent / "golden"
    spec = importlib.util.spec_from_file_location("gen_loss_goldens", g / "gen_loss_goldens.py")
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)

    seen = {}

    def ce(logits, labels):
        seen["shape"] = tuple(logits.shape)
        assert logits.is_contiguous() and labels.is_contiguous() and int(labels.min()) >= -1
        x = logits.float()
        keep = labels >= 0
        picked = x.gather(-1, labels.clamp(min=0).unsqueeze(-1)).squeeze(-1)
        return torch.where(keep, torch.logsumexp(x, -1) - picked, torch.zeros(()))   # negative label -> 0, like lp_cross_entropy

    monkeypatch.setattr(ops, "cross_entropy", ce)
    for rec in json.loads((g / "ref_loss.json").read_text()):
        logits, labels, weights = gen.make_case(rec["seed"], rec["batch"], rec["seq"], rec["vocab"], rec["frac"])
        loss, logged = megatron.loss_func(weights, labels, logits)
        assert seen["shape"] == (rec["batch"] * (rec["seq"] - 1), rec["vocab"])
        assert abs(float(loss) - rec["loss"]) <= 2e-6 * abs(rec["loss"]), (float(loss), rec["loss"])
        assert float(logged["lm loss"]) == float(loss)


def test_named_to_hf_inverts_the_kernel_layout():
    """Checkpoint export of the tensors the SFT step updates: kernel layout -> HF names is the exact inverse of load_hf."""
    import torch
    from leopard_b200.config import leopard_llava, reduced
    from leopard_b200.native_weights import load_hf, named_to_hf
    from leopard_b200.weights import SyntheticCheckpoint, tensor_specs
    cfg = reduced(leopard_llava(), vit_layers=1, dec_layers=2, vocab=324)       # padded vocabulary
    ck = SyntheticCheckpoint(cfg, seed=9)
    nw = load_hf(cfg, lambda n: ck[n], device="cpu")
    back = named_to_hf(cfg, nw.named)
    want = [n for n, _, _ in tensor_specs(cfg)]
    assert sorted(back) == sorted(want)
    for n in want:
        assert back[n].shape == ck[n].shape and torch.equal(back[n].to(torch.bfloat16), ck[n]), n
    # an in-place update of a native tensor shows up in the export (views, not copies)
    nw.named["dec.1.gateup_w"][40].fill_(3.0)          # granule 0, second half -> up_proj row 8
    assert float(named_to_hf(cfg, nw.named)["language_model.model.layers.1.mlp.up_proj.weight"][8].min()) == 3.0


def test_collate_and_train_step_shims_match_the_script():
    """megatron.mm_collate_fn against batches collated by the reference's own function (gen_collate_goldens.py), and
    megatron.train_step's plumbing (get_batch keys, None batches skipped, Megatron's return tuple)."""
    import importlib.util
    import json
    from pathlib import Path

    import torch
    from leopard_b200 import megatron

    g = Path(__file__).parent / "golden"
    spec = importlib.util.spec_from_file_location("gen_collate_goldens", g / "gen_collate_goldens.py")
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    recs = json.loads((g / "ref_collate.json").read_text())
    assert len(recs) == len(gen.CASES)
    for k, (case, want) in enumerate(zip(gen.CASES, recs)):
        got = megatron.mm_collate_fn(gen.make_batch(case, seed=k))
        if want is None:
            assert got is None
            continue
        assert set(got) == set(want)
        for key, w in want.items():
            assert list(got[key].shape) == w["shape"] and str(got[key].dtype) == w["dtype"], (k, key)
            assert got[key].float().flatten().tolist() == w["data"], (k, key)
        assert got["input_ids"].shape[1] % 8 == 0

    class Recorder:
        def step(self, tokens, labels, loss_mask, images):
            self.seen = (tokens.dtype, labels.dtype, loss_mask.dtype, tuple(tokens.shape), tuple(images.shape))
            return 1.25, 0.5

    rec = Recorder()
    batches = iter([None, megatron.mm_collate_fn(gen.make_batch(gen.CASES[0]))])
    out = megatron.train_step(None, batches, [object()], rec, None, None)
    assert out == ({"lm loss": 1.25}, 0, 0.5, None)
    assert rec.seen == (torch.long, torch.long, torch.float32, (2, 16), (3, 3, 4, 4))

    class ListRecorder:
        def step(self, tokens, labels, loss_mask, images):
            self.n = (len(tokens), len(labels), len(loss_mask), len(images), tuple(tokens[1].shape))
            return 2.0, 1.0

    lrec = ListRecorder()
    b0 = megatron.mm_collate_fn(gen.make_batch(gen.CASES[0]))
    b1 = megatron.mm_collate_fn(gen.make_batch(gen.CASES[4]))
    assert megatron.train_step(None, iter([b0, None, b1]), [object()], lrec, num_microbatches=2)[0] == {"lm loss": 2.0}
    assert lrec.n == (2, 2, 2, 2, (1, 24))


def test_checkpoint_export_reads_the_live_trained_tensors():
    """GPTModel.state_dict_for_save_checkpoint after training: the nested Megatron dict holds what is on the device
    (kernel layout inverted, Megatron's interleaves applied), not the tensors the model was loaded from."""
    import types
    import torch
    from leopard_b200 import megatron
    from leopard_b200.config import leopard_llava, reduced
    from leopard_b200.native_weights import load_hf
    from leopard_b200.weights import SyntheticCheckpoint
    cfg = reduced(leopard_llava(), vit_layers=1, dec_layers=1, vocab=324)
    ck = SyntheticCheckpoint(cfg, seed=21)
    nw = load_hf(cfg, lambda n: ck[n], device="cpu")
    model = types.SimpleNamespace(cfg=cfg, engine=types.SimpleNamespace(w=nw), _hf_getter=lambda n: ck[n])
    sd = megatron.GPTModel.state_dict_for_save_checkpoint(model)["language_model"]
    flat0 = megatron.flatten_language_model_state(sd)
    ref = megatron.hf_to_megatron(cfg, lambda n: ck[n])
    assert set(flat0) == set(ref) and all(torch.equal(flat0[k], ref[k].to(flat0[k].dtype)) for k in ref)
    # "train": bump the fused gate/up weight's first up-projection row and one projector bias in place
    nw.named["dec.0.gateup_w"][32].fill_(0.5)
    nw.named["proj.b2"][7] = 0.25
    flat1 = megatron.flatten_language_model_state(megatron.GPTModel.state_dict_for_save_checkpoint(model)["language_model"])
    h4 = flat1["encoder.layers.0.mlp.dense_h_to_4h.weight"]           # [gate ; up]
    assert float(h4[cfg.dec.ffn].float().min()) == 0.5 and torch.equal(h4[: cfg.dec.ffn], flat0["encoder.layers.0.mlp.dense_h_to_4h.weight"][: cfg.dec.ffn])
    assert float(flat1["mm_projector.encoder.linear_fc2.bias"][7]) == 0.25


def test_padded_batch_plan_reproduces_the_hf_merge_layout():
    """plan_padded_batch (leopard_b200/hf.py) against the oracle's step-by-step restatement of HF's
    `_merge_input_ids_with_image_features`: same max_len, the kept rows of every batch row land exactly where
    HF's final_attention_mask is 1, positions 0 .. S_b - 1, image features consumed in batch-row order."""
    import pytest
    from types import SimpleNamespace as NS
    from leopard_b200.hf import plan_padded_batch
    from oracle import model as om
    IMG, PAD, P, D = 7, 0, 5, 4
    cfg = NS(image_token_index=IMG)
    table = torch.arange(20 * D, dtype=torch.float32).view(20, D) + 1
    ck = NS(f32=lambda name: table)

    def check(ids, mask):
        n = int((ids == IMG).sum())
        feats = 1000 + torch.arange(n * P * D, dtype=torch.float32).view(n, P, D)
        emb, fmask, pos = om.merge_hf_batched(ck, cfg, ids, mask, feats, PAD)
        max_len, left, rows = plan_padded_batch(ids, mask, IMG, P)
        assert max_len == emb.shape[1]
        k = 0
        for b, (t0, t1, nb, S_b, off) in enumerate(rows):
            want = torch.zeros(max_len, dtype=torch.long)
            want[off:off + S_b] = 1
            assert torch.equal(fmask[b], want), (b, fmask[b], want)
            assert torch.equal(pos[b, off:off + S_b], torch.arange(S_b))
            # what the packed engine builds for this row: kept ids with every placeholder expanded to its crop's P rows
            packed = []
            for t in range(t0, t1):
                if int(ids[b, t]) == IMG:
                    packed.append(feats[k]); k += 1
                else:
                    packed.append(table[int(ids[b, t])][None])
            assert torch.equal(torch.cat(packed, 0), emb[b, off:off + S_b])
        return left

    # left padding, unequal image counts (the batched-generation case)
    ids = torch.tensor([[PAD, PAD, 3, IMG, 4, 5], [2, IMG, 3, IMG, 9, 6], [PAD, PAD, PAD, 8, 9, 11]])
    assert check(ids, (ids != PAD).long()) is True
    # right padding with the same number of images per row
    ids = torch.tensor([[3, IMG, 4, PAD, PAD], [2, IMG, 3, 9, 6]])
    assert check(ids, (ids != PAD).long()) is False
    # B = 1 without a mask reduces to the expansion the reference uses
    ids = torch.tensor([[3, IMG, 4, IMG]])
    check(ids, torch.ones_like(ids))
    # refused: right padding with unequal image counts (HF's own step 3 misplaces the rows), holes, placeholders under the mask
    ids = torch.tensor([[3, IMG, 4, PAD, PAD], [2, IMG, 3, IMG, 6]])
    with pytest.raises(NotImplementedError):
        plan_padded_batch(ids, (ids != PAD).long(), IMG, P)
    with pytest.raises(NotImplementedError):
        plan_padded_batch(torch.tensor([[1, PAD, 2, 3]]), torch.tensor([[1, 0, 1, 1]]), IMG, P)
    with pytest.raises(ValueError):
        plan_padded_batch(torch.tensor([[IMG, 2, 3]]), torch.tensor([[0, 1, 1]]), IMG, P)
