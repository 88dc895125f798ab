"""Drop-in for madeleine/utils/trainer.py: the loss glue (calculate_losses) and the training-step caller."""
import time

import numpy as np
import torch

from .. import ops
from .loss import GOT as _B200_GOT
from .utils import set_model_precision, smooth_rank_measure

DEVICE = torch.device("cuda" if torch.cuda.is_available() else "cpu")
HE_POSITION = 0
WHOLE_VIEW_POSITION = 0


def calculate_losses(STAINS, loss_fn_interMod, loss_fn_interMod_local, loss_fn_intraMod, wsi_embs, token_embs,
                     modality_labels_withoutHE, args):
    """trainer.py:20-77 — per stain: select the cases that have it, global InfoNCE (+ weighted GOT, + intra-modality
    InfoNCE on the two half views), summed.  Returns (loss, at_least_one_stain_flag); loss is -1 when nothing applies."""
    losses = []
    pending_local = []          # (slot, raw GOT loss) issued on side streams, joined once before the final sum
    atleast_two_loss_flag = False
    dev = wsi_embs["HE"].device
    # The availability mask comes from the dataloader on the CPU (as in the reference); keeping the case selection on
    # the host avoids the device sync that boolean-mask indexing of CUDA tensors costs every step.
    labels = modality_labels_withoutHE.detach().to("cpu").bool()
    counts = labels.sum(dim=0).tolist()
    for stain_idx, stain in enumerate(STAINS):
        if counts[stain_idx] <= 1:
            continue
        stain_mask = labels[:, stain_idx].nonzero(as_tuple=True)[0].to(dev, non_blocking=True)   # row indices of the cases
        if loss_fn_interMod:
            if args.global_loss != "info-nce":
                raise AssertionError("invalid global loss")
            he = wsi_embs["HE"][:, WHOLE_VIEW_POSITION, :, stain_idx][stain_mask]
            ihc = wsi_embs[stain][:, WHOLE_VIEW_POSITION, :][stain_mask]
            losses.append(loss_fn_interMod(query=he, positive_key=ihc, symmetric=args.symmetric_cl))
        if loss_fn_interMod_local:
            he_tokens = token_embs["HE"][:, :, :, stain_idx][stain_mask]
            ihc_tokens = token_embs[stain].squeeze()[stain_mask]
            if loss_fn_interMod_local is _B200_GOT and he_tokens.is_cuda:
                # the per-stain OT problems are independent and small (<= 65 CTAs each): overlap them on side streams
                slot = len(pending_local) % 4
                pending_local.append((slot, loss_fn_interMod_local(he_tokens, ihc_tokens, subsample=256, _slot=slot)))
            else:
                losses.append(loss_fn_interMod_local(he_tokens, ihc_tokens, subsample=256) * args.local_loss_weight)
        if loss_fn_intraMod:
            he1, he2 = wsi_embs["HE"][:, 1, :, stain_idx][stain_mask], wsi_embs["HE"][:, 2, :, stain_idx][stain_mask]
            st1, st2 = wsi_embs[stain][:, 1, :][stain_mask], wsi_embs[stain][:, 2, :][stain_mask]
            losses.append(loss_fn_intraMod(query=he1, positive_key=he2, symmetric=args.symmetric_cl))
            losses.append(loss_fn_intraMod(query=st1, positive_key=st2, symmetric=args.symmetric_cl))
        atleast_two_loss_flag = True
    if pending_local:
        ops.got_join(dev, sorted({slot for slot, _ in pending_local}))
        losses.extend(raw * args.local_loss_weight for _, raw in pending_local)
    if len(losses) > 0:
        loss = sum(losses)
    else:
        loss = -1
        assert loss == -1 and not atleast_two_loss_flag, "Loss should be -1 if there are no losses to calculate"
    return loss, atleast_two_loss_flag


def train_loop(args, loss_fn_interMod, loss_fn_interMod_local, loss_fn_intraMod, ssl_model, epoch, dataloader, optimizer,
               scheduler_warmup, scheduler):
    """trainer.py:80-144 — one epoch. Same control flow as the reference; the HE embeddings for the rank metric are
    gathered on the device and copied once per epoch instead of one blocking .cpu() per step."""
    n_views = 3 if loss_fn_intraMod else 1
    ssl_model.train()
    torch_precision = set_model_precision(args.precision)
    autocast_on = torch_precision in (torch.bfloat16, torch.float16)
    ep_loss = torch.zeros((), device=DEVICE)
    fb_time = 0.0
    all_embeds = []
    for b_idx, data in enumerate(dataloader):
        if epoch == 0 and b_idx == 0:
            print("Using precision:", torch_precision)
        s_fb = time.time()
        modality_labels_withoutHE = data["modality_labels"][:, HE_POSITION + 1:]
        optimizer.zero_grad()
        with torch.amp.autocast(device_type="cuda", dtype=torch_precision, enabled=autocast_on):
            wsi_embs, token_embs = ssl_model(data, device=DEVICE, n_views=n_views)
            loss, atleast_two_loss_flag = calculate_losses(args.STAINS, loss_fn_interMod, loss_fn_interMod_local, loss_fn_intraMod,
                                                           wsi_embs, token_embs, modality_labels_withoutHE, args)
        all_embeds.append(wsi_embs["HE"][:, WHOLE_VIEW_POSITION, :, 0].detach().to(torch.float32))
        if not atleast_two_loss_flag:
            print("Skipping batch with only HE")
            continue
        loss.backward()
        optimizer.step()
        if epoch <= args.warmup_epochs:
            scheduler_warmup.step()
        else:
            scheduler.step()
        if (b_idx % 3) == 0:
            print(f"Loss for batch: {b_idx} = {loss:.3f}")
        ep_loss += loss.detach()
        fb_time += time.time() - s_fb
    all_embeds_tensor = torch.cat(all_embeds, dim=0).cpu()
    rank = smooth_rank_measure(all_embeds_tensor)
    return float(ep_loss), rank
