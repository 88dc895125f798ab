#!/usr/bin/env python
"""Headline benchmark of the MADELEINE hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--precision fp32|bf16]

Metric (BASELINE.json): slides/sec (fwd+bwd) at N=2000 x D=512.  One step = BASELINE.json configs[1] per GPU:
16 cases x 2 stains (HE + one IHC) = 32 bags x 2000 patch embeddings x 512-d, forward through the drop-in
``MADELEINE.forward(train=True)`` in train mode (dropout on) + symmetric InfoNCE (tau = 0.001) via
``calculate_losses`` + ``loss.backward()`` + the fused AdamW update (so the bf16 operand planes of the weights are
re-packed every step, as in training; ``--no-optimizer`` times forward + backward alone).  For N > 1 every rank owns 16 more cases (weak scaling); the slide
embeddings are all-gathered once before the loss and parameter gradients are summed with one all-reduce.

  value    whole-job slides/s with the inputs resident in HBM
  e2e      same step through the same public call, but fed from pinned HOST memory every step (H2D inside the timed
           region) and with the loss read back to the host every step
  roofline the attention-pooling kernel (the metric's "pooling HBM GB/s vs peak"), timed live with CUDA events inside
           the timed region; `roofline_gemm` reports the tcgen05 GEMMs the same way
  cpu_baseline / --impl reference: the CPU oracle (torch-CPU restatement of the reference, oracle/) on the host cores,
           on a bounded sample of the same workload
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from argparse import Namespace

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))

CASES_PER_GPU, N_STAINS, N_TOKENS, D_IN = 16, 2, 2000, 512
MODS = ["HE", "IHC"]
TAU = 0.001
METRIC = "slides/sec (fwd+bwd) at N=2000xD=512"


def read_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """SM clock and throttle reasons sampled every ~10 ms through NVML (nvidia_ml_py) while the timed region runs;
    falls back to `nvidia-smi -lms` when NVML is not importable."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self.power_w, self.power_limit_w = [], None
        self._stop = threading.Event()
        self._thread = None
        self._smi = None

    def _visible_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.index])
            except (ValueError, IndexError):
                return self.index
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._visible_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            try:
                self.power_limit_w = pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1e3
            except Exception:  # noqa: BLE001
                pass

            def poll():
                while not self._stop.is_set():
                    try:
                        self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        for bit, name in self.REASONS.items():
                            if mask & bit:
                                self.reasons.add(name)
                        self.power_w.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)
                    except Exception:  # noqa: BLE001
                        pass
                    time.sleep(0.01)

            self._thread = threading.Thread(target=poll, daemon=True)
            self._thread.start()
        except Exception:  # noqa: BLE001
            self._smi = _SmiSampler(self.index)
            self._smi.start()

    def stop(self):
        if self._smi is not None:
            return self._smi.stop()
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1.0)
        out = {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples), "source": "nvml, 10 ms period"}
        if self.power_w:
            # NVML's board power is a slow average (~1 s): over a 0.1 s timed region it mostly reflects the warm-up steps before it.
            # Sustained, the step draws 983 W of the 1000 W limit at ~1515 MHz (tools/ab_probe.py --power, profiles/r02_s2_ab_probes.json)
            out["power_w"] = round(statistics.median(self.power_w), 1)
            out["power_limit_w"] = self.power_limit_w
        return out


class _SmiSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 50"}


def model_cfg(precision):
    return Namespace(MODALITIES=MODS, wsi_encoder="abmil", patch_embedding_dim=D_IN, wsi_encoder_hidden_dim=512,
                     activation="softmax", n_heads=4, b200_precision=precision)


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """The reference's own CPU path (oracle port) on the host cores; rank 0 only."""
    if rank != 0:
        return
    import oracle
    from weights import make_state_dict, make_feats
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cases = CASES_PER_GPU                                         # the GPU arm's own batch: 16 cases x 2 stains x 2000 tokens per step
    sd = {k: v.clone().requires_grad_(True) for k, v in make_state_dict(0, n_mod=2).items()}
    feats = make_feats(1, cases, N_STAINS, N_TOKENS, D_IN)
    labels = torch.ones(cases, 1)

    def step():
        for v in sd.values():
            v.grad = None
        embs, toks = oracle.madeleine_forward_train(sd, feats, MODS)
        loss, _ = oracle.calculate_losses(MODS[1:], embs, toks, labels, temperature=TAU, symmetric=True)
        loss.backward()
        return float(loss)

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    steps = max(1, min(args.steps, 5))                            # bounded: ~2 s per step on 16 host cores
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    value = cases * N_STAINS / dt
    sample = f"{cases} cases x {N_STAINS} stains x {N_TOKENS} tokens per step, {steps} steps, oracle port (torch CPU, fp32)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "slides/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1] at the metric's fixed N=2000: 16 cases x 2 stains per step (the GPU arm's batch; bounded number of steps)",
                   "bags_per_step": cases * N_STAINS, "bags_per_gpu": cases * N_STAINS,
                   "tokens_per_bag": N_TOKENS, "d_in": D_IN, "loss": "symmetric InfoNCE tau=0.001"},
        "cpu_baseline": {"value": value, "unit": "slides/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "slides/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline_quick():
    import oracle
    from weights import make_state_dict, make_feats
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cases = CASES_PER_GPU
    sd = {k: v.clone().requires_grad_(True) for k, v in make_state_dict(0, n_mod=2).items()}
    feats = make_feats(1, cases, N_STAINS, N_TOKENS, D_IN)
    labels = torch.ones(cases, 1)
    times = []
    for it in range(3):
        for v in sd.values():
            v.grad = None
        t0 = time.perf_counter()
        embs, toks = oracle.madeleine_forward_train(sd, feats, MODS)
        loss, _ = oracle.calculate_losses(MODS[1:], embs, toks, labels, temperature=TAU, symmetric=True)
        loss.backward()
        times.append(time.perf_counter() - t0)
    dt = statistics.median(times[1:])
    return {"value": cases * N_STAINS / dt, "unit": "slides/s", "cores": cores, "kind": "port",
            "sample": f"{cases} cases x {N_STAINS} stains x {N_TOKENS} tokens (the bench batch), fwd+bwd, median of 2 after 1 warm-up (oracle port, torch CPU fp32)"}


def canonical_block(precision, dev, steps):
    """The reference's shipped pre-training configuration (scripts/launch_pretrain_withStainEncodings.sh; the one BASELINE.md's
    only published number — 38.4 cases/s on 3 x 3090 Ti — is quoted on): batch 65 cases x 5 stains x 2048 tokens, stain
    encodings, ACROBAT availability, InfoNCE (tau 0.001) + Graph-OT, fwd+bwd+AdamW, train mode; token window off and on."""
    from madeleine.models.Model import MADELEINE
    from madeleine.utils.loss import InfoNCE, GOT
    from madeleine.utils.trainer import calculate_losses
    from madeleine_b200.optim import FusedAdamW
    from weights import make_state_dict
    mods = ["HE", "HER2", "PGR", "KI67", "ER"]
    bs, T = 65, 2048
    gen = torch.Generator().manual_seed(0)
    labels = (torch.rand(bs, 5, generator=gen) < torch.tensor([1.0, 0.46, 0.73, 0.73, 0.73])).float()
    labels[:, 0] = 1
    feats = torch.randn(bs, 5, T, D_IN, device=dev) * labels.to(dev)[:, :, None, None]
    largs = Namespace(global_loss="info-nce", symmetric_cl=True, local_loss_weight=1.0)
    loss_fn = InfoNCE(temperature=TAU)
    res = {"workload": "65 cases x 5 stains x 2048 x 512, stain encodings, 28 % of the stain slots missing, InfoNCE + GOT, "
                       "fwd+bwd+AdamW, train mode", "published_reference_cases_per_s_3x3090Ti": 38.4}
    # the run's precision with the token window off and on, then the precision the reference's scripts ship (--precision bfloat16:
    # the published 38.4 cases/s on 3 x 3090Ti were measured under bf16 autocast)
    variants = [(precision, "off"), (precision, "batch")] + ([("bf16", "batch")] if precision != "bf16" else [])
    for prec, window in variants:
        cfg = Namespace(MODALITIES=mods, wsi_encoder="abmil", patch_embedding_dim=D_IN, wsi_encoder_hidden_dim=512, activation="softmax",
                        n_heads=4, b200_precision=prec, b200_token_window=window)
        model = MADELEINE(cfg, stain_encoding=True)
        model.load_state_dict(make_state_dict(3, n_mod=5, stain_encoding=True))
        model.to(dev).train()
        opt = FusedAdamW(model.parameters(), lr=1e-4)

        def step():
            model.zero_grad(set_to_none=True)
            embs, toks = model({"feats": feats, "modality_labels": labels}, device=dev, n_views=1)
            loss, _ = calculate_losses(mods[1:], loss_fn, GOT, None, embs, toks, labels[:, 1:], largs)
            loss.backward()
            opt.step()
            return loss

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        key = f"token_window_{window}" if prec == precision else f"shipped_precision_{prec}_token_window_{window}"
        res[key] = {"ms_per_step": ms, "cases_per_s": bs / (ms * 1e-3), "slides_per_s": float(labels.sum()) / (ms * 1e-3),
                    "loss": float(loss.detach())}
        del model, opt
        torch.cuda.empty_cache()
    res["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 1e9
    return res


def ragged_block(precision, dev, steps):
    """BASELINE configs[1] as written (N = 1 only): 16 cases x 2 stains with N_i ~ randint(200, 4001) per bag (generator seed 1234,
    SURVEY.md §8d), packed back to back ([sum N_i, 512] + cu_seqlens, no padding) through ``forward_packed`` — token embeddings
    included, as forward(train=True) computes them — + symmetric InfoNCE (tau = 0.001) + backward + fused AdamW, train mode."""
    from madeleine.models.Model import MADELEINE
    from madeleine.utils.loss import InfoNCE
    from madeleine_b200.optim import FusedAdamW
    from weights import make_state_dict
    model = MADELEINE(model_cfg(precision), stain_encoding=False)
    model.load_state_dict(make_state_dict(0, n_mod=2), strict=True)
    model.to(dev).train()
    opt = FusedAdamW(model.parameters(), lr=1e-4)
    g = torch.Generator().manual_seed(1234)
    lens = torch.randint(200, 4001, (CASES_PER_GPU * N_STAINS,), generator=g)
    cu = torch.zeros(lens.numel() + 1, dtype=torch.int32)
    cu[1:] = lens.cumsum(0)
    total = int(cu[-1])
    x = torch.randn(total, D_IN, device=dev)
    cu_dev = cu.to(dev)
    loss_fn = InfoNCE(temperature=TAU)

    def step():
        model.zero_grad(set_to_none=True)
        slide, _tokens = model.forward_packed(x, cu_dev, want_tokens=True)      # bags 2c / 2c + 1 = HE / IHC slide of case c
        loss = loss_fn(query=slide[0::2], positive_key=slide[1::2], symmetric=True)
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"workload": "16 cases x 2 stains, N_i ~ randint(200, 4001) (seed 1234), bag-packed, fwd+bwd+AdamW, train mode",
            "bags": int(lens.numel()), "tokens": total, "min_len": int(lens.min()), "max_len": int(lens.max()), "ms_per_step": ms,
            "slides_per_s": lens.numel() / (ms * 1e-3), "tokens_per_s": total / (ms * 1e-3),
            "padded_tokens_if_batched_dense": int(lens.max()) * int(lens.numel()), "loss": float(loss.detach())}


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp32_fwd", "bf16"])
    ap.add_argument("--eval-mode", action="store_true", help="model.eval(): dropout off")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--timeline", action="store_true", help="N > 1: add a per-phase device timeline of the step (max and mean over ranks)")
    ap.add_argument("--no-sustained", action="store_true", help="skip the ~3 s sustained-state run")
    ap.add_argument("--no-canonical", action="store_true", help="skip the two extra N = 1 blocks: configs[1] with ragged bags and the reference's canonical 65 x 5 x 2048 configuration")
    ap.add_argument("--no-optimizer", action="store_true",
                    help="time forward + backward only; by default every timed step also runs the fused AdamW update, so the "
                         "kernel-layout copies of the weights are re-packed every step as in real training (nothing is cached "
                         "across steps)")
    ap.add_argument("--with-optimizer", action="store_true", help="(default; kept for older command lines)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from madeleine.models.Model import MADELEINE
    from madeleine.utils.loss import InfoNCE
    from madeleine.utils.trainer import calculate_losses
    from madeleine_b200 import _lib, parallel
    from weights import make_state_dict

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(1234)
    model = MADELEINE(model_cfg(args.precision), stain_encoding=False)
    model.load_state_dict(make_state_dict(0, n_mod=2), strict=True)
    model.to(dev)
    model.eval() if args.eval_mode else model.train()
    loss_fn = InfoNCE(temperature=TAU)
    largs = Namespace(global_loss="info-nce", symmetric_cl=True, local_loss_weight=1.0)

    B = CASES_PER_GPU
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    feats_dev = torch.randn(B, N_STAINS, N_TOKENS, D_IN, generator=g, device=dev)          # 131 MB > 126 MB L2
    from madeleine_b200.utils import hostmem
    saved_affinity = os.sched_getaffinity(0)
    hostmem.bind_to_gpu(local_rank)     # first-touch the pinned pages from the GPU's NUMA node (no-op on single-node hosts)
    feats_host = []                     # e2e: pinned host buffers
    for _ in range(2):
        buf = hostmem.pinned_empty((B, N_STAINS, N_TOKENS, D_IN))
        buf.normal_()
        feats_host.append(buf)
    os.sched_setaffinity(0, saved_affinity)
    labels = torch.ones(B, N_STAINS)
    labels_dev = labels.to(dev)
    labels_global = torch.ones(B * world, N_STAINS)      # the loader knows the whole batch's availability mask (case list)
    parallel.enable_gradient_sync(world > 1)

    optimizer = None
    args.with_optimizer = not args.no_optimizer
    if args.with_optimizer:
        from madeleine_b200.optim import FusedAdamW
        optimizer = FusedAdamW(model.parameters(), lr=1e-4)       # reference: optim.AdamW(lr=args.lr), lr 1e-4 in the scripts

    def step(feats, ctx=None, update=True):
        lab, lab_dev, lab_global = ctx or (labels, labels_dev, labels_global)
        model.zero_grad(set_to_none=True)
        embs, toks = model({"feats": feats}, device=dev, n_views=1)
        # the availability mask stays on the host (as the reference's dataloader delivers it)
        if world > 1:
            embs, lab = parallel.gather_slide_embeddings(embs, lab_dev, global_labels_host=lab_global)
        loss, ok = calculate_losses(MODS[1:], loss_fn, None, None, embs, toks, lab[:, 1:], largs)
        loss.backward()               # world > 1: the encoder backward all-reduces its flat gradient buffer (enable_gradient_sync)
        if optimizer is not None and update:
            optimizer.step()
        return loss

    def sharded_parity():
        """N > 1 only, once, before anything is timed: the sharded step (cases split over the ranks, ONE all-gather of the
        slide embeddings, gradients summed by the encoder's all-reduce) against the same global batch evaluated by rank 0
        alone — loss and parameter gradients.  Dropout off (model.eval()) so that both evaluations are deterministic."""
        was_training = model.training
        model.eval()
        loss_sh = step(feats_dev, update=False).detach().clone()
        named = [(n, p) for n, p in model.named_parameters() if p.grad is not None]
        g_sh = torch.cat([p.grad.reshape(-1) for _, p in named]).clone()
        feats_all = torch.empty((world * B,) + tuple(feats_dev.shape[1:]), device=dev)
        dist.all_gather_into_tensor(feats_all, feats_dev)
        out = None
        if rank == 0:
            parallel.enable_gradient_sync(False)
            model.zero_grad(set_to_none=True)
            embs, toks = model({"feats": feats_all}, device=dev, n_views=1)
            loss_1, _ = calculate_losses(MODS[1:], loss_fn, None, None, embs, toks, labels_global[:, 1:], largs)
            loss_1.backward()
            g_1 = torch.cat([p.grad.reshape(-1) for _, p in named])
            worst, worst_name, o = 0.0, "", 0
            total = float(g_1.double().norm())
            for n, p in named:
                k = p.numel()
                ref = g_1[o:o + k].double()
                err = float((g_sh[o:o + k].double() - ref).norm()) / max(float(ref.norm()), 1e-6 * total)
                if err > worst:
                    worst, worst_name = err, n
                o += k
            out = {"what": f"sharded step on {world} ranks vs the same global batch ({world * B} cases) on rank 0 alone, eval mode",
                   "loss_sharded": float(loss_sh), "loss_single": float(loss_1.detach()),
                   "loss_rel": abs(float(loss_sh) - float(loss_1.detach())) / max(abs(float(loss_1.detach())), 1e-12),
                   "grad_rel": float((g_sh.double() - g_1.double()).norm()) / total,
                   "grad_rel_worst_param": worst, "worst_param": worst_name}
            parallel.enable_gradient_sync(True)
            model.zero_grad(set_to_none=True)
        del feats_all
        dist.barrier()
        model.train(was_training)
        return out

    parity = sharded_parity() if world > 1 else None
    if world > 1:
        # The parity check keeps every GPU busy for a while (the ranks > 0 spin in NCCL's barrier while rank 0 evaluates the global
        # batch), and the step runs at the board's power limit: without a pause the K timed steps of an N > 1 run would start in
        # the throttled state while an N = 1 run starts from an idle GPU.  Same protocol for every N: idle, W warm-up, K timed.
        torch.cuda.synchronize()
        time.sleep(2.0)
        dist.barrier()

    def timed(n_steps, feats_fn, read_loss, ctx=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_steps):
            loss = step(feats_fn(i), ctx)
            if read_loss:
                loss.item()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / n_steps

    for _ in range(max(args.warmup, 3)):
        step(feats_dev)
    torch.cuda.synchronize()

    # ---- device-resident timing, with live per-kernel events for the roofline ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # Inside the timed region only the pooling launches (the metric's named kernel: `roofline`) are bracketed with CUDA events; the
    # GEMM launches are timed in a second pass of K steps right after it (`roofline_gemm`, `kernel_ms_per_step`): 24 more event
    # records per step would sit between kernels that otherwise chain through programmatic dependent launch.
    _lib.start_timing({"mdl_pool_fwd", "mdl_pool_weights", "mdl_pool_bwd_dlogit"})
    _lib.launch_count[0] = 0
    _lib.native_launches(reset=True)
    ms_step = timed(args.steps, lambda i: feats_dev, read_loss=False)
    launches = (_lib.launch_count[0] + _lib.native_launches()) // args.steps
    clocks = sampler.stop() if rank == 0 else None
    kt = _lib.stop_timing()
    _lib.start_timing({"mdl_gemm_nt", "mdl_gemm_gated", "mdl_gemm_tn_accum"})
    timed(args.steps, lambda i: feats_dev, read_loss=False)
    kt.update(_lib.stop_timing())

    bags_per_step = B * N_STAINS * world
    value = bags_per_step / (ms_step * 1e-3)

    # ---- optional: where does a multi-GPU step spend its time?  Events at the phase boundaries, K more steps ----
    timeline = None
    if args.timeline:
        from madeleine_b200 import ops as _ops

        def mark(tl, label):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            tl.append((label, ev))

        per_step = []
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for _ in range(args.steps):
            tl = []
            _ops.timeline = tl
            mark(tl, "step_begin")
            model.zero_grad(set_to_none=True)
            embs, toks = model({"feats": feats_dev}, device=dev, n_views=1)
            mark(tl, "forward_done")
            lab = labels
            if world > 1:
                embs, lab = parallel.gather_slide_embeddings(embs, labels_dev, global_labels_host=labels_global)
            mark(tl, "allgather_done")
            loss, _ = calculate_losses(MODS[1:], loss_fn, None, None, embs, toks, lab[:, 1:], largs)
            mark(tl, "loss_done")
            loss.backward()
            mark(tl, "backward_done")
            if optimizer is not None:
                optimizer.step()
            mark(tl, "optimizer_done")
            _ops.timeline = None
            per_step.append(tl)
        torch.cuda.synchronize()
        labels_seq = [lbl for lbl, _ in per_step[0]]
        spans = torch.zeros(len(labels_seq) - 1, device=dev)
        for tl in per_step:
            for i in range(len(tl) - 1):
                spans[i] += tl[i][1].elapsed_time(tl[i + 1][1])
        spans /= len(per_step)
        mx, mean = spans.clone(), spans.clone()
        if world > 1:
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(mean, op=dist.ReduceOp.SUM)
            mean /= world
        timeline = {"spans_ms": {f"{labels_seq[i]} -> {labels_seq[i + 1]}": {"mean_over_ranks": float(mean[i]), "max_over_ranks": float(mx[i])}
                                 for i in range(len(labels_seq) - 1)},
                    "note": "device time between events recorded at the phase boundaries of a step (average of K steps); inside "
                            "backward: bwd_begin -> bwd_kernels_done covers the encoder's backward kernels with the early (90 %) "
                            "gradient all-reduce in flight on NCCL's stream, then the remaining 2 MB all-reduce, then the join"}

    # ---- BASELINE configs[3]: batch = 64 cases sharded over 8 ranks = 8 cases per rank (timed for any N > 1) ----
    config3 = None
    if world > 1:
        b3 = 8
        ctx3 = (torch.ones(b3, N_STAINS), torch.ones(b3, N_STAINS, device=dev), torch.ones(b3 * world, N_STAINS))
        feats3 = feats_dev[:b3].contiguous()
        for _ in range(3):
            step(feats3, ctx3)
        ms3 = timed(args.steps, lambda i: feats3, read_loss=False, ctx=ctx3)
        config3 = {"workload": f"BASELINE configs[3]: {b3} cases x {N_STAINS} stains x {N_TOKENS} per rank, global batch {b3 * world} cases "
                               "sharded, one all-gather of slide embeddings before InfoNCE", "cases_per_gpu": b3, "global_batch_cases": b3 * world,
                   "ms_per_step": ms3, "slides_per_s": b3 * N_STAINS * world / (ms3 * 1e-3)}

    # ---- end to end: every step's features come from pinned HOST memory (H2D inside the timed region, staged one batch
    # ahead on a copy stream by DevicePrefetcher) and every step's loss is read back to the host (async D2H into a pinned
    # buffer, consumed one step later so the host keeps one step of launches queued) ----
    e2e = None
    if not args.no_e2e:
        from madeleine_b200.utils.prefetch import DevicePrefetcher

        def run_e2e(n_steps):
            batches = ({"feats": feats_host[i % 2]} for i in range(n_steps))
            LAG = 2                                    # the host reads the loss of step i-2 while step i is being queued (tools/e2e_probe.py)
            host_loss = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(LAG + 1)]
            events = [torch.cuda.Event() for _ in range(LAG + 1)]
            seen = []
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i, batch in enumerate(DevicePrefetcher(batches, dev)):
                loss = step(batch["feats"])
                host_loss[i % (LAG + 1)].copy_(loss.detach(), non_blocking=True)
                events[i % (LAG + 1)].record()
                if i >= LAG:
                    events[(i - LAG) % (LAG + 1)].synchronize()
                    seen.append(float(host_loss[(i - LAG) % (LAG + 1)]))
            for j in range(max(0, n_steps - LAG), n_steps):
                events[j % (LAG + 1)].synchronize()
                seen.append(float(host_loss[j % (LAG + 1)]))
            e1.record()
            torch.cuda.synchronize()
            assert len(seen) == n_steps and all(x == x for x in seen)
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.barrier()
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms) / n_steps

        run_e2e(5)
        e2e_runs = [run_e2e(args.steps) for _ in range(3)]      # K steps each; the median is reported, all three are listed
        ms_e2e = sorted(e2e_runs)[1]
        # what the link alone gives on this box (the e2e number is bounded by it on hosts with slow pinned copies)
        scratch = torch.empty_like(feats_dev)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()              # all ranks copy at the same time: the rate a rank sees when the host serves N GPUs
        torch.cuda.synchronize()
        c0.record()
        for _ in range(4):
            scratch.copy_(feats_host[0], non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        h2d_gbs = 4 * feats_host[0].numel() * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        del scratch
        h2d_ms = feats_host[0].numel() * 4 / (h2d_gbs * 1e9) * 1e3
        bound = ("host h2d: one 131 MB batch takes %.2f ms at the %.1f GB/s a rank gets while all %d ranks copy, the step itself %.2f ms"
                 % (h2d_ms, h2d_gbs, world, ms_step)) if h2d_ms > 0.9 * ms_step else "device (the copy hides under the step)"
        e2e = {"value": bags_per_step / (ms_e2e * 1e-3), "unit": "slides/s", "ms_per_step": ms_e2e, "bound": bound,
               "h2d_bytes_per_step": feats_host[0].numel() * 4, "d2h_bytes_per_step": 4,
               "h2d_gbs_alone": h2d_gbs, "runs_ms_per_step": e2e_runs,
               "protocol": "5 warm-up steps, then 3 x K timed steps back to back; median.  Every run pays its own pipeline fill (the first "
                           "batch's 131 MB copy is not overlapped: 2.4 ms / K per step) and runs in the sustained power state",
               "note": "pinned host features staged one batch ahead on a copy stream; every step's loss is read back inside the "
                       "timed region, two steps deferred so the host stays ahead of the device (tools/e2e_probe.py); "
                       "h2d_gbs_alone = this box's pinned H2D rate for one batch with the GPU otherwise idle (N > 1: all ranks copying "
                       "at the same time)"}

    # ---- the same device-resident step in the SUSTAINED state (~3 s of continuous stepping): the step runs at the board's power
    # limit, so the K steps timed right after the warm-up are a few per cent faster than the steady state.
    # (All ranks: timed() is a collective.) ----
    sustained = None
    if not args.no_sustained:
        n_sus = max(args.steps, int(3000.0 / ms_step))
        sus = ClockSampler(local_rank)
        if rank == 0:
            sus.start()
        ms_sus = timed(n_sus, lambda i: feats_dev, read_loss=False)
        sus_clk = sus.stop() if rank == 0 else None
        if rank == 0:
            sustained = {"steps": n_sus, "ms_per_step": ms_sus, "slides_per_s": bags_per_step / (ms_sus * 1e-3),
                         "sm_mhz": sus_clk.get("sm_mhz"), "power_w": sus_clk.get("power_w"),
                         "power_limit_w": sus_clk.get("power_limit_w"), "reasons": sus_clk.get("reasons")}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = read_peaks()
    npl = 1 if args.precision == "bf16" else 2
    bags_local = B * N_STAINS
    # algorithmic bytes of ONE pooling launch: per bag N*2048*s + N*4*4 + 2048*4, s = 2*nplanes
    pool_bytes = bags_local * (N_TOKENS * 2048 * 2 * npl + N_TOKENS * 4 * 4 + 2048 * 4)
    pool_ms = statistics.mean(kt["mdl_pool_fwd"])
    pool_gbs = pool_bytes / (pool_ms * 1e-3) / 1e9
    # DRAM traffic of the same launch from the committed `ncu --set full` capture (profiles/), fp32-mode workload only
    traffic, traffic_src = None, None
    ncu_json = os.path.join(REPO, "profiles", "r02_pool_fwd_ncu.json")
    if args.precision == "fp32" and os.path.exists(ncu_json):
        cap = json.load(open(ncu_json))
        if cap.get("algorithmic_bytes") == pool_bytes:
            traffic, traffic_src = cap["traffic_bytes"], "profiles/r02_pool_fwd_ncu.json (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full)"
    roofline = {"kernel": "pool_fwd_kernel (attention pooling, forward; the softmax weights come from the separate pool_weights_kernel)", "bound": "hbm",
                "achieved": pool_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": pool_gbs / peaks["hbm_gbs"],
                "traffic": traffic, "traffic_source": traffic_src, "avg_launch_ms": pool_ms,
                "algorithmic_bytes_per_launch": pool_bytes, "peak_source": peaks["source"]}
    # tcgen05 GEMMs: algorithmic FLOPs per step (fp32-equivalent, 1x; the 3-pass split issues 3x this on the bf16 pipe)
    tokens = bags_local * N_TOKENS
    flops_fwd = tokens * (2 * D_IN * 512 + 2 * 512 * 512 + 2 * 512 * 2048 + 4 * 2 * (2 * 512 * 512) + 2 * 2048 * 128)
    flops_bwd = tokens * (2 * (2 * 512 * 2048 + 2 * 512 * 512 + 4 * 2 * (2 * 512 * 512)) + 2 * D_IN * 512)
    gemm_ms = sum(sum(kt.get(k, [])) for k in ("mdl_gemm_nt", "mdl_gemm_gated", "mdl_gemm_tn_accum")) / args.steps
    issued = {"fp32": 3.0, "bf16": 1.0, "fp32_fwd": (3 * flops_fwd + flops_bwd) / (flops_fwd + flops_bwd)}[args.precision]
    tf = (flops_fwd + flops_bwd) / (gemm_ms * 1e-3) / 1e12
    roofline_gemm = {"kernel": "gemm_tcgen05_kernel (all forward/dgrad/wgrad GEMMs of a step)", "bound": "tensor", "achieved": tf,
                     "achieved_bf16_issue": tf * issued, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                     "frac": tf / peaks["bf16_tflops_sustained"], "frac_bf16_issue": tf * issued / peaks["bf16_tflops_sustained"],
                     "ms_per_step": gemm_ms, "note": "algorithmic fp32-equivalent FLOPs; the 3-pass split-bf16 mode issues 3x on the tensor pipe; timed in a "
                     "second pass of K steps right after the timed region"}
    pool_bwd_ms = statistics.mean(kt["mdl_pool_bwd_dlogit"]) if kt.get("mdl_pool_bwd_dlogit") else None
    roofline["pool_weights_avg_launch_ms"] = statistics.mean(kt["mdl_pool_weights"]) if kt.get("mdl_pool_weights") else None

    out = {
        "metric": METRIC, "value": value, "unit": "slides/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32 (3-pass split-bf16 on tcgen05, fp32 accumulate)", "bf16": "bf16 (tcgen05, fp32 accumulate)",
                  "fp32_fwd": "f32-grade forward (3-pass split-bf16), bf16 backward GEMMs (1 pass), fp32 accumulate"}[args.precision],
        "data": "synthetic",
        "config": {"workload": "BASELINE configs[1] at the metric's fixed N=2000: 16 cases x 2 stains per GPU, symmetric InfoNCE tau=0.001, "
                               "MADELEINE.forward(train=True) + calculate_losses + backward, train mode (dropout on)"
                               + (" + fused AdamW step (weights re-packed every step)" if args.with_optimizer else " (no optimiser step)")
                               if not args.eval_mode else "same, eval mode",
                   "bags_per_gpu": bags_local, "tokens_per_bag": N_TOKENS, "d_in": D_IN, "parallelism": f"dp{world} (cases sharded, 1 all-gather of slide embeddings + grad all-reduce, 90 % of it overlapped with the backward pass)",
                   "l2": "inputs larger than L2 (131 MB features + >1 GB activations per step, L2 = 126 MB)"},
        "clocks": clocks, "gpu_launches": launches, "e2e": e2e, "roofline": roofline, "roofline_gemm": roofline_gemm,
        "kernel_ms_per_step": {k: sum(v) / args.steps for k, v in kt.items()}, "pool_bwd_avg_launch_ms": pool_bwd_ms,
    }
    # whole step against the tensor roofline: algorithmic (1x, fp32-equivalent) FLOPs of everything a step computes per
    # second of step time, over the measured sustained bf16 rate (46.2 GFLOP/bag fwd+bwd with the token
    # projector's backward; this step has no local loss, so 44.0 GFLOP/bag)
    step_tf = (flops_fwd + flops_bwd) * world / (ms_step * 1e-3) / 1e12
    out["roofline_step"] = {"bound": "tensor", "achieved": step_tf / world, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                            "frac": step_tf / world / peaks["bf16_tflops_sustained"], "per_gpu": True,
                            "algorithmic_gflop_per_bag": (flops_fwd + flops_bwd) / bags_local / 1e9,
                            "frac_bf16_issue": step_tf * issued / world / peaks["bf16_tflops_sustained"],
                            "note": "algorithmic FLOPs x bags/s over the sustained bf16 rate; x3 issued in the fp32-grade mode"}
    out["sustained"] = sustained
    out["parity"] = parity
    if world > 1:
        out["small_collectives"] = parallel.PeerExchange.status()
    if timeline is not None:
        out["timeline"] = timeline
    if config3 is not None:
        out["config3"] = config3
    if world == 1 and not args.no_canonical:
        del feats_dev, feats_host
        torch.cuda.empty_cache()
        out["ragged"] = ragged_block(args.precision, dev, args.steps)
        out["canonical"] = canonical_block(args.precision, dev, 5)
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline_quick()
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
