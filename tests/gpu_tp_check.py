"""Multi-GPU check of the tensor-parallel decoder (run under torchrun, one rank per GPU, 2 / 4 / 8 ranks):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/gpu_tp_check.py
  1. all-gather fused into the GEMM (lp_gemm_bf16_tp) == NCCL all_gather + lp_gemm_bf16, bit for bit;
  2. GEMM with the reduce-scatter fused into its epilogue + lp_tp_reduce vs the sum of the ranks' bf16 partial products;
  3. the whole TP + sequence-parallel decoder (comm = "fused" and "nccl") against the TP = 1 engine on the same weights;
  4. timing of one decoder layer at the bench shape (S = 4796 / 4800 rows): fused kernels vs the NCCL sequencing.
Prints one JSON line per item on rank 0; exit code 0 only if every check passed on every rank."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from leopard_b200 import ops, tp as TP  # noqa: E402
from leopard_b200.config import leopard_llava, reduced  # noqa: E402
from leopard_b200.engine import PrefillEngine  # noqa: E402
from leopard_b200.native_weights import load_hf  # noqa: E402
from leopard_b200.tp_fused import FusedComm  # noqa: E402
from leopard_b200.weights import SyntheticCheckpoint  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True

    def say(**kw):
        if rank == 0:
            print(json.dumps(kw), flush=True)

    def all_ok(flag):
        t = torch.tensor([1 if flag else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    H = 4096
    S = 4796 if 4796 % world == 0 else 4800
    sl = S // world
    fc = FusedComm(S, H, device=dev)
    g = torch.Generator(device=dev).manual_seed(1234)            # same seed on every rank: identical "global" tensors
    x_full = (torch.randn((S, H), device=dev, generator=g) * 0.5).to(torch.bfloat16)

    # ---- 1. all-gather + GEMM
    gw = torch.Generator(device=dev).manual_seed(100 + rank)
    w_col = (torch.randn((3072, H), device=dev, generator=gw) * H ** -0.5).to(torch.bfloat16)      # this rank's column shard
    for rep in range(3):
        fc.local_rows().copy_(x_full[rank * sl:(rank + 1) * sl])
        got = fc.allgather_gemm(w_col)
        gathered = torch.empty((S, H), dtype=torch.bfloat16, device=dev)
        dist.all_gather_into_tensor(gathered, x_full[rank * sl:(rank + 1) * sl].contiguous())
        want = ops.gemm(gathered, w_col, block_n=256)
        same = torch.equal(got, want) and torch.equal(gathered, x_full)
        ok &= all_ok(same)
        torch.cuda.synchronize()
        dist.barrier()
    say(check="allgather_gemm == nccl all_gather + gemm (bitwise, 3 repeats)", ok=ok)

    # ---- 2. GEMM + reduce-scatter
    w_row = (torch.randn((H, 2048), device=dev, generator=gw) * 2048 ** -0.5).to(torch.bfloat16)     # this rank's K shard
    a_loc = (torch.randn((S, 2048), device=dev, generator=gw) * 0.5).to(torch.bfloat16)
    resid = x_full[rank * sl:(rank + 1) * sl].contiguous()
    got = fc.gemm_reduce_scatter(a_loc, w_row, resid)
    partial = ops.gemm(a_loc, w_row, block_n=256)                                                  # bf16 partial sums [S, H]
    parts = [torch.empty_like(partial) for _ in range(world)]
    dist.all_gather(parts, partial)
    tot = sum(p[rank * sl:(rank + 1) * sl].float() for p in parts).to(torch.bfloat16).float() + resid.float()
    err = float((got.float() - tot).abs().max())
    ok2 = all_ok(err <= 2.0 ** -6 * float(tot.abs().max()))
    ok &= ok2
    say(check="gemm_reduce_scatter == residual + sum of bf16 partials", ok=ok2, max_abs=err)

    # ---- 3. whole decoder, reduced depth, full widths
    cfg = reduced(leopard_llava(), vit_layers=1, dec_layers=2, vocab=2048)
    ck = SyntheticCheckpoint(cfg, seed=1234)
    W = TP.shard_decoder_weights(cfg, lambda n: ck[n], world, rank, dev)
    S3 = 256 * world
    emb = (torch.randn((S3, H), device=dev, generator=torch.Generator(device=dev).manual_seed(7)) * 0.5).to(torch.bfloat16)
    fc3 = FusedComm(S3, H, device=dev)
    sl3 = S3 // world
    logits_f = TP.TPDecoder(cfg, W, comm="fused", fused=fc3).forward(emb[rank * sl3:(rank + 1) * sl3].clone(), S3, last_only=False)
    logits_n = TP.TPDecoder(cfg, W, comm="nccl").forward(emb[rank * sl3:(rank + 1) * sl3].clone(), S3, last_only=False)
    eng = PrefillEngine(cfg, load_hf(cfg, lambda n: ck[n], dev))
    ref, _ = eng.decoder(emb.clone(), torch.arange(S3, dtype=torch.int32, device=dev), [0, S3], "all", torch.float32)
    e_fn = float((logits_f - logits_n).abs().max() / ref.abs().max())
    e_f1 = float((logits_f - ref).abs().max() / ref.abs().max())
    ok3 = all_ok(e_fn < 1e-2 and e_f1 < 2e-2)
    ok &= ok3
    say(check="TP decoder (2 layers): fused vs nccl, fused vs TP=1 engine (relative)", ok=ok3, fused_vs_nccl=e_fn, fused_vs_tp1=e_f1)

    # ---- 4. one full-size layer: fused vs NCCL sequencing
    cfg1 = reduced(leopard_llava(), vit_layers=1, dec_layers=1, vocab=2048)
    W1 = TP.shard_decoder_weights(cfg1, lambda n: SyntheticCheckpoint(cfg1, seed=5)[n], world, rank, dev)
    x_loc = x_full[rank * sl:(rank + 1) * sl].clone()

    def timed(dec, n=6):
        for _ in range(2):
            dec.forward(x_loc.clone(), S)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            dec.forward(x_loc.clone(), S)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t_f = timed(TP.TPDecoder(cfg1, W1, comm="fused", fused=fc))
    t_n = timed(TP.TPDecoder(cfg1, W1, comm="nccl"))
    say(check="one decoder layer + final norm + lm_head, S rows", S=S, tp=world, fused_ms=round(t_f, 3), nccl_ms=round(t_n, 3),
        speedup=round(t_n / t_f, 3))
    say(result="TP CHECK OK" if ok else "TP CHECK FAILED")
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
