"""world_size-2 gloo test (CPU) of the sharding helpers: one all-gather of slide embeddings + summed gradients must
reproduce the single-process loss and gradients of the full batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _loss(embs, labels):
    """Full-batch symmetric InfoNCE on the (gathered) embeddings via the oracle (CPU)."""
    import oracle
    he = embs["HE"][:, 0, :, 0]
    ihc = embs["IHC"][:, 0, :]
    keep = labels[:, 1].bool()
    return oracle.info_nce(he[keep], ihc[keep], temperature=0.1, symmetric=True)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from madeleine_b200 import parallel
        torch.manual_seed(0)
        B = 4
        W = torch.randn(16, 8)                                  # a shared "model"
        x_all = torch.randn(world * B, 2, 16)
        labels_all = torch.ones(world * B, 2)
        labels_all[1, 1] = 0
        lin = torch.nn.Linear(16, 8, bias=False)
        with torch.no_grad():
            lin.weight.copy_(W.t())
        x = x_all[rank * B:(rank + 1) * B]
        out = lin(x)                                            # [B, 2, 8]
        embs = {"HE": out[:, 0:1, :].unsqueeze(-1).expand(-1, -1, -1, 1), "IHC": out[:, 1:2, :]}
        g_embs, g_labels = parallel.gather_slide_embeddings(embs, labels_all[rank * B:(rank + 1) * B])
        assert g_embs["HE"].shape == (world * B, 1, 8, 1) and g_embs["IHC"].shape == (world * B, 1, 8)
        assert torch.equal(g_labels, labels_all)
        loss = _loss(g_embs, g_labels)
        loss.backward()
        parallel.allreduce_gradients(lin)
        # single-process reference on the full batch
        ref = torch.nn.Linear(16, 8, bias=False)
        with torch.no_grad():
            ref.weight.copy_(W.t())
        o = ref(x_all)
        ref_loss = _loss({"HE": o[:, 0:1, :].unsqueeze(-1), "IHC": o[:, 1:2, :]}, labels_all)
        ref_loss.backward()
        torch.testing.assert_close(loss.detach(), ref_loss.detach(), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(lin.weight.grad, ref.weight.grad, rtol=1e-4, atol=1e-6)
        # the row-addressed path: the encoder's slide-embedding matrix itself is gathered (rank-major blocks of
        # [view][case][modality] rows) and the per-modality tensors are rebuilt from it on demand
        from madeleine_b200 import ops
        for n_views in (1, 3):
            lin.weight.grad = None
            base = lin(torch.randn(world * n_views * B * 2, 16, generator=torch.Generator().manual_seed(5))[rank * n_views * B * 2:
                                                                                                       (rank + 1) * n_views * B * 2])
            local = ops.EmbeddingDict()
            blk = base.view(n_views, B, 2, 8)
            local["HE"] = blk[:, :, 0].permute(1, 0, 2).unsqueeze(-1)
            local["IHC"] = blk[:, :, 1].permute(1, 0, 2)
            local.b200_set(base, B, 2, n_views)
            g2, lab2 = parallel.gather_slide_embeddings(local, labels_all[rank * B:(rank + 1) * B], global_labels_host=labels_all)
            assert lab2 is labels_all and g2.b200_world == world and g2.b200_base.shape == (world * n_views * B * 2, 8)
            # same tensors as gathering the per-modality entries one by one
            plain, _ = parallel.gather_slide_embeddings(dict(local), labels_all[rank * B:(rank + 1) * B])
            torch.testing.assert_close(g2["HE"], plain["HE"])
            torch.testing.assert_close(g2["IHC"], plain["IHC"])
            for case in (0, B - 1, B, world * B - 1):
                for mod in (0, 1):
                    for view in range(n_views):
                        row = g2.b200_row(case, mod, view)
                        want = (g2["HE"][case, view, :, 0] if mod == 0 else g2["IHC"][case, view])
                        assert torch.equal(g2.b200_base[row], want)
            g2.b200_base.square().sum().backward()          # all-gather backward = this rank's slice
            assert lin.weight.grad is not None
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


def test_gather_and_gradient_sum_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def test_single_process_passthrough():
    from madeleine_b200 import parallel
    x = torch.randn(3, 4)
    assert parallel.all_gather_rows(x) is x
    embs = {"HE": torch.randn(3, 1, 8, 1), "IHC": torch.randn(3, 1, 8)}
    lab = torch.ones(3, 2)
    e2, l2 = parallel.gather_slide_embeddings(embs, lab)
    assert e2 is embs and l2 is lab
