"""torchrun worker for tests/test_gpu_syncbn.py::test_two_nccl_ranks_*: a 2-rank data-parallel train step (SyncBN through
parallel.PeerSums AND through torch.distributed, gradients through parallel.GradBucket) against the single-device step on
the concatenated batch, computed by rank 0 on its own GPU.  Writes a JSON report (rank 0)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import cases as C  # noqa: E402
import tcb_oracle as O  # noqa: E402
from cvpr2021_vspw_implement_b200 import engine as E  # noqa: E402
from cvpr2021_vspw_implement_b200 import parallel as P  # noqa: E402

T, H, W, N_LOC = 3, 65, 97, 2


def main(out_path):
    world, rank, local = P.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    imgs, labs = O.synthetic_clip(T, world * N_LOC, H, W, C.NUM_CLASS, seed=41, block=16, ignore_frac=0.0)
    report = {}
    peer = P.PeerSums()
    # exchange self-test: integers are exact in fp64, the total must be identical on every rank
    t = (torch.arange(5000, device=dev, dtype=torch.float64) + 1) * (rank + 1)
    peer.all_reduce_sums(t)
    want = (torch.arange(5000, device=dev, dtype=torch.float64) + 1) * sum(r + 1 for r in range(world))
    assert torch.equal(t, want), "PeerSums total differs"
    for mode, group, prec in (("peer/fp32", peer, "fp32"), ("peer/bf16x3", peer, "bf16x3"), ("nccl/bf16x3", E.TorchDistGroup(), "bf16x3")):
        E.set_precision(prec)
        # ---- data-parallel step -----------------------------------------------------------------------------------
        m = C.no_dropout(C.build("Clip_PSP", "resnet50dilated", 31).to(dev).train())
        P.broadcast_parameters(m)
        bucket = P.GradBucket(m.parameters())
        E.set_syncbn(True, group=group)
        sl = slice(rank * N_LOC, (rank + 1) * N_LOC)
        bucket.zero_grad()
        loss, acc = m(C.feed([i[sl] for i in imgs], [l[sl] for l in labs], True, dev))
        loss.backward()
        bucket.all_reduce_mean()
        loss_dp = float(P.mean_scalar(loss).item())
        if mode == "peer/fp32":
            # a second step with the same bucket: now the gradient chunks are all-reduced DURING the backward pass
            # (GradBucket overlap); same weights and inputs, so the averaged gradients must repeat
            first = bucket.flat[:bucket.numel].clone()
            assert bucket.last_overlapped == 0
            state_first = {k: v.clone() for k, v in m.state_dict().items()}
            bucket.zero_grad()
            loss2, _ = m(C.feed([i[sl] for i in imgs], [l[sl] for l in labs], True, dev))
            loss2.backward()
            bucket.all_reduce_mean()
            assert bucket.last_overlapped == len(bucket._chunks), (bucket.last_overlapped, len(bucket._chunks))
            again = bucket.flat[:bucket.numel]
            rel = float((again.double() - first.double()).norm() / first.double().norm())
            assert rel < 1e-5, f"overlapped all-reduce changed the gradients: rel {rel:.3e}"
            report["overlap_rel"] = rel
            bucket.flat[:bucket.numel].copy_(first)
            m.load_state_dict(state_first)  # (running statistics of ONE step are what the single-device run is compared with)
        E.set_syncbn(False)
        # ---- single-device global batch (rank 0) --------------------------------------------------------------------
        if rank == 0:
            ref = C.no_dropout(C.build("Clip_PSP", "resnet50dilated", 31).to(dev).train())
            E.set_grad_sink(None)
            lr, _ = ref(C.feed(imgs, labs, True, dev))
            lr.backward()
            torch.cuda.synchronize()
            worst, worst_bn = (0.0, ""), 0.0
            for (k, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
                if q.grad is None:
                    assert p.grad is None, k
                    continue
                rn = float(q.grad.double().norm())
                if rn < 1e-7:
                    continue
                e = float((p.grad.double() - q.grad.double()).norm() / rn)
                if e > worst[0]:
                    worst = (e, k)
                if k.endswith(("bn1.weight", "bn1.bias", "bn3.weight", "bn3.bias", ".1.weight", ".1.bias")):
                    worst_bn = max(worst_bn, e)
            sd, sr = m.state_dict(), ref.state_dict()
            wr = max(C.rel_err(sd[k].cpu(), sr[k].cpu()) for k in sd if k.endswith(("running_mean", "running_var")))
            report[mode] = {"loss_rel": abs(loss_dp - lr.item()) / abs(lr.item()), "worst_grad_rel_l2": worst[0], "worst_grad": worst[1],
                            "bn_affine_worst": worst_bn, "worst_running": wr}
        dist.barrier()
    peer.close()
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(report, f)
        print(json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
