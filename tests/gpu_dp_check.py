"""Multi-GPU check (run under torchrun, one rank per GPU): rank 0 owns a queue of independent requests, scatters them
cost-balanced over NCCL, every rank prefills its share on the native path, results are gathered and compared with the
CPU oracle on rank 0.  No data-path collective: only the batch scatter / result gather (leopard_b200/dp.py)."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from leopard_b200 import dp  # noqa: E402
from leopard_b200.config import leopard_llava, reduced  # noqa: E402
from leopard_b200.engine import PrefillEngine  # noqa: E402
from leopard_b200.native_weights import load_hf  # noqa: E402
from leopard_b200.weights import SyntheticCheckpoint, synth_inputs  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = reduced(leopard_llava(), vit_layers=2, dec_layers=2, vocab=2048)
    ck = SyntheticCheckpoint(cfg, seed=1234)
    eng = PrefillEngine(cfg, load_hf(cfg, lambda n: ck[n], dev))
    reqs = None
    shapes = [(3, 10), (1, 30), (0, 25), (2, 8), (1, 12), (4, 6)]
    if rank == 0:
        reqs = []
        for k, (n, t) in enumerate(shapes):
            pix, ids = synth_inputs(cfg, n_crops=n, n_text=t, seed=10 + k)
            reqs.append((ids[0], pix if n else None))
    local_reqs, mine = dp.scatter_requests(reqs, dev, src=0)
    logits, _, _ = eng.prefill(local_reqs, "last") if local_reqs else (torch.zeros(0, cfg.dec.vocab, device=dev), None, None)
    res = dp.gather_results(logits.float(), mine, len(shapes))
    ok = True
    if rank == 0:
        from oracle import model as om
        for k, (ids, pix) in enumerate(reqs):
            if pix is None:
                e = ck.f32("language_model.model.embed_tokens.weight")[ids]
                ref = om.decoder_forward(ck, cfg, e, torch.arange(ids.numel()), last_only=True)[0]
            else:
                ref = om.prefill_logits(ck, cfg, ids[None], pix.to(torch.bfloat16).float(), last_only=True)[0, 0]
            err = float((res[k].cpu() - ref).abs().max() / ref.abs().max())
            print(f"request {k}: crops {shapes[k][0]} owner-balanced, rel err vs oracle {err:.4f}")
            ok &= err < 3e-2
        print("assignment rank0:", mine, "DP CHECK", "OK" if ok else "FAILED")
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
