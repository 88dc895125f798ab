"""The chunked fp64 yardstick of the BASELINE-size GPU parity tests equals the oracle's monolithic evaluation (CPU, small)."""
import torch

import oracle
import parity_utils as pu
from weights import make_state_dict


def test_chunked_oracle_step_matches_monolithic():
    mods = ["HE", "ER", "PR"]
    bs, T = 5, 24
    sd_cpu = make_state_dict(2, n_mod=3, stain_encoding=True)
    g = torch.Generator().manual_seed(3)
    labels = torch.ones(bs, 3)
    labels[1, 2] = 0
    feats = (torch.randn(bs, 3, T, 512, generator=g) * labels[:, :, None, None]).double()
    sd_a = pu.to_oracle_sd(sd_cpu, "cpu")
    loss_a, embs_a, _ = pu.oracle_train_step(sd_a, feats, mods, labels[:, 1:], stain_encoding=True, temperature=0.001,
                                             use_local=True, loss_seed=4, chunk_rows=4)
    sd_b = pu.to_oracle_sd(sd_cpu, "cpu")
    embs_b, toks_b = oracle.madeleine_forward_train(sd_b, feats, mods, stain_encoding=True)
    torch.manual_seed(4)
    loss_b, _ = oracle.calculate_losses(mods[1:], embs_b, toks_b, labels[:, 1:], temperature=0.001, symmetric=True, use_local=True)
    loss_b.backward()
    torch.testing.assert_close(loss_a, loss_b.detach(), rtol=1e-10, atol=1e-10)
    for m in mods:
        torch.testing.assert_close(embs_a[m], embs_b[m].detach(), rtol=1e-10, atol=1e-12)
    for k in sd_a:
        ga, gb = sd_a[k].grad, sd_b[k].grad
        assert (ga is None) == (gb is None), k
        if ga is not None:
            assert float((ga - gb).norm()) <= 1e-8 * float(gb.norm()) + 1e-12, k


def test_packed_infonce_step_and_rank_statistics():
    sd_cpu = make_state_dict(1, n_mod=2)
    lens = [5, 9, 3, 7]
    cu = [0, 5, 14, 17, 24]
    x = torch.randn(24, 512, generator=torch.Generator().manual_seed(0)).double()
    sd = pu.to_oracle_sd(sd_cpu, "cpu")
    loss, emb = pu.oracle_packed_infonce_step(sd, x, cu, 0.1)
    ref = oracle.encode_packed({k: v.detach() for k, v in sd.items()}, x, cu)
    torch.testing.assert_close(emb, ref)
    torch.testing.assert_close(loss, oracle.info_nce(ref[:2], ref[2:], temperature=0.1, symmetric=True))
    assert sd["projector.weight"].grad is not None and sd["token_projector.weight"].grad is None
    raw = torch.randn(2, 50, 1, 4)
    st = pu.rank_statistics(raw, raw)
    assert st["identical_rank_fraction"] == 1.0 and st["mismatched_positions"] == 0 and st["top8_identical"]
    raw2 = raw.clone()
    raw2[0, :, 0, 0] = raw[0, :, 0, 0].flip(0)
    assert pu.rank_statistics(raw2, raw)["identical_rank_fraction"] < 1.0
