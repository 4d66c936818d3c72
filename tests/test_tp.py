"""CPU: tensor-parallel + sequence-parallel decoder (leopard_b200/tp.py).
  * TP-sharded Megatron checkpoint split / merge follow the converter's per-key rules and are exact inverses;
  * the TP layer loop (world 2, gloo, native ops replaced by their contract restatements) reproduces the TP = 1 oracle."""
import os
from dataclasses import replace

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from leopard_b200 import tp as TP
from leopard_b200.config import leopard_llava, reduced
from leopard_b200.native_weights import hf_to_megatron, megatron_to_hf
from leopard_b200.weights import SyntheticCheckpoint


def _small_cfg():
    cfg = reduced(leopard_llava(), vit_layers=1, dec_layers=2, vocab=512)
    return replace(cfg, dec=replace(cfg.dec, hidden=512, heads=8, kv_heads=4, ffn=1024), vit=replace(cfg.vit, hidden=144, heads=2, ffn=256))


@pytest.mark.parametrize("tp", [2, 4])
def test_megatron_checkpoint_shards_follow_the_converter_rules(tp):
    cfg = _small_cfg()
    ck = SyntheticCheckpoint(cfg, seed=3)
    lm = hf_to_megatron(cfg, lambda n: ck[n])
    shards = TP.shard_megatron_state(lm, tp)
    d = cfg.dec
    k = "encoder.layers.0."
    # column-parallel: dim 0; row-parallel: dim 1 (hf2megatron_llava.py:1176-1184)
    assert shards[0][k + "self_attention.query_key_value.weight"].shape == ((d.heads + 2 * d.kv_heads) * d.head_dim // tp, d.hidden)
    assert shards[1][k + "self_attention.dense.weight"].shape == (d.hidden, d.hidden // tp)
    assert shards[1][k + "mlp.dense_4h_to_h.weight"].shape == (d.hidden, d.ffn // tp)
    assert torch.equal(shards[0][k + "input_norm.weight"], lm[k + "input_norm.weight"])                      # replicated
    # every shard's dense_h_to_4h is [gate chunk ; up chunk] (hf2megatron_llava.py:1287-1295)
    gate = ck[f"language_model.model.layers.0.mlp.gate_proj.weight"]
    up = ck[f"language_model.model.layers.0.mlp.up_proj.weight"]
    fl = d.ffn // tp
    for r in range(tp):
        h4 = shards[r][k + "mlp.dense_h_to_4h.weight"]
        assert torch.equal(h4[:fl], gate[r * fl:(r + 1) * fl]) and torch.equal(h4[fl:], up[r * fl:(r + 1) * fl])
    # a QKV shard holds whole kv groups: [rep q heads | k | v] per group (transformer.py:713-738)
    rep, hd = d.heads // d.kv_heads, d.head_dim
    q = ck["language_model.model.layers.0.self_attn.q_proj.weight"]
    qkv1 = shards[1][k + "self_attention.query_key_value.weight"]
    g0 = d.kv_heads // tp                                                 # first kv group of rank 1
    assert torch.equal(qkv1[: rep * hd], q[g0 * rep * hd:(g0 + 1) * rep * hd])
    merged = TP.merge_megatron_shards(shards)
    assert merged.keys() == lm.keys() and all(torch.equal(merged[n], lm[n]) for n in lm)
    back = megatron_to_hf(cfg, merged)
    assert torch.equal(back["language_model.model.layers.1.mlp.up_proj.weight"], ck["language_model.model.layers.1.mlp.up_proj.weight"])


def _tp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import native_ops
        TP.ops = native_ops                                              # contract restatements instead of the CUDA kernels
        cfg = _small_cfg()
        ck = SyntheticCheckpoint(cfg, seed=5)
        W = TP.shard_decoder_weights(cfg, lambda n: ck[n], world, rank, "cpu")
        S = 48
        g = torch.Generator().manual_seed(1)
        x = (torch.randn((S, cfg.dec.hidden), generator=g) * 0.5).to(torch.bfloat16)
        sl = S // world
        logits = TP.TPDecoder(cfg, W, comm="nccl").forward(x[rank * sl:(rank + 1) * sl].clone(), S, last_only=False)
        q.put((rank, logits))
    finally:
        dist.destroy_process_group()


def test_tp2_sequence_parallel_decoder_matches_the_tp1_oracle_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33000 + os.getpid() % 2000
    procs = [ctx.Process(target=_tp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.equal(got[0][1], got[1][1])                             # every rank ends with the same gathered logits
    from oracle import model as om
    cfg = _small_cfg()
    ck = SyntheticCheckpoint(cfg, seed=5)
    g = torch.Generator().manual_seed(1)
    x = (torch.randn((48, cfg.dec.hidden), generator=g) * 0.5).to(torch.bfloat16)
    ref = om.decoder_forward(ck, cfg, x.float(), torch.arange(48))
    err = float((got[0][1] - ref).abs().max() / ref.abs().max())
    assert got[0][1].shape == ref.shape and err < 3e-2, err
