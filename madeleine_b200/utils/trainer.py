"""Drop-in for madeleine/utils/trainer.py: the loss glue (``calculate_losses``) and the training-step caller
(``train_loop``).  Signatures, return values, printed lines and control flow are the reference's (trainer.py:20-144);
the internals are organised around what the device needs:

* which cases carry a stain is decided on the HOST from the availability mask the loader delivers (one small index
  upload per stain) — boolean-mask indexing of CUDA tensors would cost a device sync per stain and step;
* the per-stain Graph-OT problems are independent and small, so they are issued on side streams and joined once;
* the epoch loss and the H&E embeddings for the rank metric stay on the device until the epoch ends.
"""
import time

import torch

from .. import ops
from .loss import GOT as _B200_GOT, InfoNCE as _B200_InfoNCE
from .utils import set_model_precision, smooth_rank_measure

DEVICE = torch.device("cuda" if torch.cuda.is_available() else "cpu")
HE_POSITION = 0            # the H&E slide is modality 0 of every case
WHOLE_VIEW_POSITION = 0    # view 0 = all tokens; views 1, 2 = the two random halves (n_views = 3)
_GOT_STREAMS = 4


def _stains_with_pairs(stains, availability, device):
    """Yield (column, stain name, device row indices of the cases that have the stain) for every stain present in at least
    two cases of the batch (trainer.py:24-26: `stain_mask.sum() > 1`)."""
    present = availability.detach().to("cpu").bool()
    per_stain = present.sum(dim=0).tolist()
    for col, name in enumerate(stains):
        if per_stain[col] > 1:
            yield col, name, present[:, col].nonzero(as_tuple=True)[0].to(device, non_blocking=True)


def _slide_pair(wsi_embs, stain, col, view, rows):
    """(H&E embedding paired with this stain, stain embedding) of one view for the selected cases."""
    return wsi_embs["HE"][:, view, :, col][rows], wsi_embs[stain][:, view, :][rows]


_ROW_PLANS = {}


def _global_terms_by_rows(STAINS, loss_fn_interMod, loss_fn_intraMod, wsi_embs, availability, args):
    """All slide-level InfoNCE terms of a step as ONE kernel-side sum over row index lists of the encoder's slide-embedding
    matrix (ops.InfoNCERowsFn).  Returns None when the fast path does not apply (foreign loss objects, foreign embeddings)."""
    base = getattr(wsi_embs, "b200_base", None)
    fns = [f for f in (loss_fn_interMod, loss_fn_intraMod) if f]
    if base is None or not base.is_cuda or not fns or any(type(f) is not _B200_InfoNCE or f.reduction != "mean" for f in fns):
        return None
    if loss_fn_interMod and args.global_loss != "info-nce":
        raise AssertionError("invalid global loss")
    if loss_fn_intraMod and wsi_embs.b200_n_views != 3:
        return None
    present = availability.detach().to("cpu").bool()
    key = (present.numpy().tobytes(), tuple(present.shape), tuple(STAINS), wsi_embs.b200_bs, wsi_embs.b200_n_mod, wsi_embs.b200_n_views,
           wsi_embs.b200_world, bool(loss_fn_interMod), bool(loss_fn_intraMod), str(base.device))
    plan = _ROW_PLANS.get(key)
    if plan is None:
        counts = present.sum(dim=0).tolist()
        pairs, kinds = [], []
        for col in range(len(STAINS)):
            if counts[col] <= 1:
                continue
            cases = present[:, col].nonzero(as_tuple=True)[0]
            he = lambda v: wsi_embs.b200_row(cases, HE_POSITION, v)          # noqa: E731
            st = lambda v: wsi_embs.b200_row(cases, col + 1, v)              # noqa: E731
            if loss_fn_interMod:
                pairs.append((he(WHOLE_VIEW_POSITION), st(WHOLE_VIEW_POSITION)))
                kinds.append("inter")
            if loss_fn_intraMod:
                pairs += [(he(1), he(2)), (st(1), st(2))]
                kinds += ["intra", "intra"]
        flat, terms, o = [], [], 0
        for q_rows, k_rows in pairs:
            m = q_rows.numel()
            terms.append((o, o + m, m))
            flat += [q_rows.to(torch.int32), k_rows.to(torch.int32)]
            o += 2 * m
        idx = torch.cat(flat).pin_memory().to(base.device, non_blocking=True) if flat else None
        plan = (idx, terms, kinds)
        if len(_ROW_PLANS) > 256:
            _ROW_PLANS.clear()
        _ROW_PLANS[key] = plan
    idx, terms, kinds = plan
    if not terms:
        return None
    fn_of = {"inter": loss_fn_interMod, "intra": loss_fn_intraMod}
    full = [(qo, ko, m, fn_of[k].temperature, args.symmetric_cl) for (qo, ko, m), k in zip(terms, kinds)]
    with torch.cuda.device(base.device):
        return ops.InfoNCERowsFn.apply(base, {"idx": idx, "terms": full})


def calculate_losses(STAINS, loss_fn_interMod, loss_fn_interMod_local, loss_fn_intraMod, wsi_embs, token_embs,
                     modality_labels_withoutHE, args):
    """Sum over the stains of: global InfoNCE(H&E, stain) + local_loss_weight x GOT(H&E tokens, stain tokens) + the two
    intra-modality InfoNCE terms on the half views, each evaluated on the cases that have the stain.
    Returns ``(loss, flag)``; ``flag`` says whether any stain contributed, ``loss`` is -1 when none did."""
    base = getattr(wsi_embs, "b200_base", None)
    device = base.device if base is not None else wsi_embs["HE"].device
    terms, local_terms = [], []            # local_terms: (side-stream slot, unweighted GOT loss)
    any_stain = False
    fused_global = _global_terms_by_rows(STAINS, loss_fn_interMod, loss_fn_intraMod, wsi_embs, modality_labels_withoutHE, args)
    if fused_global is not None:
        terms.append(fused_global)
    per_stain = () if (fused_global is not None and not loss_fn_interMod_local) else \
        _stains_with_pairs(STAINS, modality_labels_withoutHE, device)          # nothing left to do stain by stain
    for col, stain, rows in per_stain:
        any_stain = True
        if loss_fn_interMod and fused_global is None:
            if args.global_loss != "info-nce":
                raise AssertionError("invalid global loss")
            q, k = _slide_pair(wsi_embs, stain, col, WHOLE_VIEW_POSITION, rows)
            terms.append(loss_fn_interMod(query=q, positive_key=k, symmetric=args.symmetric_cl))
        if loss_fn_interMod_local:
            tok_he = token_embs["HE"][:, :, :, col][rows]
            tok_stain = token_embs[stain].squeeze()[rows]
            if loss_fn_interMod_local is _B200_GOT and tok_he.is_cuda:
                slot = len(local_terms) % _GOT_STREAMS
                local_terms.append((slot, loss_fn_interMod_local(tok_he, tok_stain, subsample=256, _slot=slot)))
            else:                           # a user-supplied local loss: plain call, as in the reference
                terms.append(loss_fn_interMod_local(tok_he, tok_stain, subsample=256) * args.local_loss_weight)
        if loss_fn_intraMod and fused_global is None:
            he_a, st_a = _slide_pair(wsi_embs, stain, col, 1, rows)
            he_b, st_b = _slide_pair(wsi_embs, stain, col, 2, rows)
            terms.append(loss_fn_intraMod(query=he_a, positive_key=he_b, symmetric=args.symmetric_cl))
            terms.append(loss_fn_intraMod(query=st_a, positive_key=st_b, symmetric=args.symmetric_cl))
    if fused_global is not None:
        any_stain = True
    if local_terms:
        ops.got_join(device, sorted({slot for slot, _ in local_terms}))
        terms += [value * args.local_loss_weight for _, value in local_terms]
    if not terms:
        assert not any_stain, "Loss should be -1 if there are no losses to calculate"
        return -1, any_stain
    return (terms[0] if len(terms) == 1 else sum(terms)), any_stain


def train_loop(args, loss_fn_interMod, loss_fn_interMod_local, loss_fn_intraMod, ssl_model, epoch, dataloader, optimizer,
               scheduler_warmup, scheduler):
    """One epoch (trainer.py:80-144): forward + losses under autocast, backward, optimiser and scheduler step per batch;
    batches without a second stain are skipped.  Returns ``(epoch loss, smooth rank of the H&E embeddings)``."""
    ssl_model.train()
    n_views = 3 if loss_fn_intraMod else 1
    dtype = set_model_precision(args.precision)
    use_autocast = dtype in (torch.bfloat16, torch.float16)
    running = torch.zeros((), device=DEVICE)          # stays on the device: no .item() per step
    he_embeddings = []
    busy = 0.0
    for step, batch in enumerate(dataloader):
        if epoch == 0 and step == 0:
            print("Using precision:", dtype)
        t0 = time.time()
        availability = batch["modality_labels"][:, HE_POSITION + 1:]
        optimizer.zero_grad()
        with torch.amp.autocast(device_type="cuda", dtype=dtype, enabled=use_autocast):
            wsi_embs, token_embs = ssl_model(batch, device=DEVICE, n_views=n_views)
            loss, usable = calculate_losses(args.STAINS, loss_fn_interMod, loss_fn_interMod_local, loss_fn_intraMod,
                                            wsi_embs, token_embs, availability, args)
        he_embeddings.append(wsi_embs["HE"][:, WHOLE_VIEW_POSITION, :, 0].detach().float())
        if not usable:
            print("Skipping batch with only HE")
            continue
        loss.backward()
        optimizer.step()
        (scheduler_warmup if epoch <= args.warmup_epochs else scheduler).step()
        if step % 3 == 0:
            print(f"Loss for batch: {step} = {loss:.3f}")
        running += loss.detach()
        busy += time.time() - t0
    rank = smooth_rank_measure(torch.cat(he_embeddings, dim=0))          # on the device: Gram matrix + eigenvalues, one scalar back
    return float(running), rank
