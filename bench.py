#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native Transformer4SED hot path.

Metric (BASELINE.json): 10 s-clips/s of a MAT-SED base training step (fused STFT->mel front end + PaSST encoder +
TransformerXL context net + frame classifier, forward + losses + backward + gradient all-reduce + AdamW) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--precision bf16|tf32|tf32x3]
                    [--workload matsed|matsed_finetune2|pmam|dasm]

`--workload` selects the other BASELINE.json configurations (default: the headline MAT-SED base step):
  matsed_finetune2  the real mean-teacher step of config/mat-sed/base/finetune2.yaml: student fwd+bwd, EMA teacher forward without
                    gradient through the sliding-window fusion (encoder_win=True, win_param=[512, 49]), fused six-loss kernel, AdamW, EMA
  pmam              PMAM post-pre-training (config/pmam/post_pretrain.yaml: PaSST_CNN, LoRA, CNN branch, d=384 decoder), 32 clips/GPU
  dasm              DASM open-vocabulary path (K = 407 queries), 8 clips/GPU

N > 1 is launched by torchrun (one rank per GPU, NCCL).  Rank 0 prints ONE JSON line.
`--impl reference` times the reference's own CPU path (the oracle port of it: the reference is Python and does not travel to
the GPU box) on the host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

BASE_KW = dict(passt_feature_layer=10, f_pool="mean_pool", decode_ratio=10, at_adapter=True, decoder="transformerXL",
               decoder_layer_num=3, decoder_pos_emd_len=1000, mlm=False)  # config/mat-sed/base/finetune2.yaml:53-62
N_SAMPLES = 320000       # 10 s @ 32 kHz (the model asserts 1000 frames: reference passt_sed.py:260)
FWD_GFLOP_PER_CLIP = 297.3
# config/mat-sed/base/finetune2.yaml:86-101 (opt.param_groups): stepped encoder LR (last 4 blocks and the norms at 2x), decoder, head
FINETUNE2_OPT = dict(encoder=dict(lr=5.0e-6, weight_decay=1.0e-4, freeze_layer=0, step_lr=4), decoder=dict(lr=1.0e-4, weight_decay=1.0e-4),
                     head=dict(lr=1.0e-4, weight_decay=1.0e-4))
# config/pmam/post_pretrain.yaml:48-79 and config/detect_any_sound (init kwargs as the golden generator uses them)
PMAM_PASST = dict(passt_feature_layer=10, class_num=30, f_pool="attention", decode_ratio=10, at_adapter=True, decoder="transformerXL",
                  decoder_layer_num=3, decoder_pos_emd_len=1000, decoder_dim=384, mlm=True,
                  lora_config=dict(r=8, lora_alpha=1, requires_grad_pretrain=False),
                  mlm_dict=dict(strategy="block", block_width=10, mask_rate=0.8, out_dim=768, mask_style=[0.9, 0.05, 0.05]))
WORKLOAD_BATCH = dict(matsed=64, matsed_finetune2=64, pmam=32, dasm=8)
MIX = (16, 6, 21, 21)    # strong, synthetic, weak, unlabeled of a 64-clip DESED batch (finetune1.yaml:12 ratio 3:1:4:4)


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:  # noqa: BLE001
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_labels(batch, seed):
    from transformer4sed_b200.utils import synth
    y = synth.synth_strong_labels(batch, 10, 1000, seed)
    return y, (y.sum(-1) > 0).float()


def sub_batches(batch):
    s = [round(batch * m / sum(MIX)) for m in MIX]
    n_strong = max(1, s[0] + s[1])
    n_weak = max(1, min(batch - n_strong, s[2])) if batch > 1 else 0
    return n_strong, n_weak


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
# ------------------------------------------------------------------------------------------------------------------
# workloads: (model, front end, parameter arena, train_step(wav) -> loss, description, host clips)
# ------------------------------------------------------------------------------------------------------------------
def build_workload(name, precision, B, dev, rank):
    import copy
    from transformer4sed_b200 import functional as F, schema
    from transformer4sed_b200.training import MeanTeacher, ParamArena, mean_teacher_step, passt_param_groups
    from transformer4sed_b200.utils import synth
    n_distinct = min(B, 8)
    wav_host = synth.synth_wav(n_distinct, N_SAMPLES, seed=1234 + rank).repeat((B + n_distinct - 1) // n_distinct, 1)[:B].contiguous()
    shadow = precision == "bf16"
    if name in ("matsed", "matsed_finetune2"):
        from transformer4sed_b200.src_models.passt.passt_sed import PaSST_SED
        net = PaSST_SED(load_pretrained_model=False, **BASE_KW)
        net.load_state_dict(synth.synth_state_dict_like(net, 4), strict=True)   # identical on every rank
        net = net.to(dev).train()
        ext = net.get_feature_extractor().eval()
        arena = ParamArena(net, passt_param_groups(net, FINETUNE2_OPT), shadow_bf16=shadow)
        y, yw = make_labels(B, 99 + rank)
        y, yw = y.to(dev), yw.to(dev)
        n_s, n_w = sub_batches(B)
        if name == "matsed":
            def train_step(wav):
                mel = ext.logmel(wav)
                strong, weak, other = net(mel)
                loss = F.bce_loss(strong[:n_s], y[:n_s])
                if n_w:
                    loss = loss + 0.5 * F.bce_loss(weak[n_s:n_s + n_w], yw[n_s:n_s + n_w])
                loss = loss + 2.0 * F.bce_loss(other["at_out"][:n_s + n_w], yw[:n_s + n_w])
                loss.backward()
                arena.step()
                return loss
            desc = ("MAT-SED base (config/mat-sed/base/finetune2.yaml init_kwargs, all 100.95 M params trainable), "
                    f"DESED-shape synthetic batch={B}/GPU of 10 s @ 32 kHz clips (320000 samples -> 1000 frames), step = fused "
                    "STFT+mel front end + student fwd + BCE strong/weak/AT losses + bwd + grad all-reduce + fused AdamW")
            return net, ext, arena, train_step, desc, wav_host
        teacher = MeanTeacher(net, arena)
        teacher.teacher.train()
        stu_kw = dict(encoder_win=False, win_param=[512, 49], mix_rate=0.5, temp_w=1)     # finetune2.yaml:64-78
        tch_kw = dict(encoder_win=True, win_param=[512, 49], mix_rate=0.5, temp_w=1)
        weights = dict(w_weak=0.5, w_at=1.0, w_cons=40.0, w_weak_cons=1.0)
        state = {"step": 1}
        if B < 3:
            raise SystemExit("matsed_finetune2 needs at least 3 clips per GPU (strong / weak / unlabelled rows)")
        n_w = max(1, min(n_w, B - n_s - 1))

        def train_step(wav):
            mel = ext.logmel(wav)
            state["step"] += 1
            total, _ = mean_teacher_step(net, teacher, arena, mel, mel, y, yw, (0, n_s), (n_s, n_s + n_w), stu_kwargs=stu_kw, tch_kwargs=tch_kw,
                                         loss_weights=weights, step_num=state["step"], ema_factor=0.999)
            return total
        desc = ("MAT-SED base mean-teacher step (config/mat-sed/base/finetune2.yaml; recipes/desed/finetune/train.py:129-199): front end, student "
                f"fwd+bwd, EMA teacher forward without gradient with encoder_win=True win_param=[512, 49] (1 global + 11 window passes folded into one "
                f"batched pass), fused six-loss kernel, grad all-reduce + fused AdamW, one-kernel EMA; batch={B}/GPU of 10 s @ 32 kHz clips")
        return net, ext, arena, train_step, desc, wav_host
    generic = lambda net: [dict(name="all", params=[p for p in net.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-4)]  # noqa: E731
    if name == "pmam":
        from transformer4sed_b200.src_models.cnn_transformer.passt_cnn import PaSST_CNN
        cnn = dict(n_in_channel=1, activation="cg", conv_dropout=0.5, kernel_size=[3] * 10, padding=[1] * 10, stride=[1] * 10,
                   nb_filters=list(schema.PMAM_FILTERS), pooling=[list(p) for p in schema.PMAM_POOLING])
        net = PaSST_CNN(dict(PMAM_PASST, load_pretrained_model=False), cnn)
        net.load_state_dict(synth.synth_state_dict_like(net, 1))
        net = net.to(dev).train()
        ext = net.get_feature_extractor().eval()
        arena = ParamArena(net, generic(net), shadow_bf16=shadow)
        g = torch.Generator().manual_seed(7 + rank)
        protos = torch.nn.functional.normalize(torch.randn(30, 768, generator=g), dim=-1).to(dev)
        labels = (torch.rand(B, 1000, 30, generator=g) < 0.1).float().to(dev)
        lw = labels.amax(1)

        def train_step(wav):
            mel = ext.logmel(wav)
            pred, other = net(mel)
            strong = F.prototype_predict(pred, protos)
            rows = other["mask_id_seq"].reshape(-1)
            loss = F.bce_loss(strong.reshape(-1, 30)[rows], labels.reshape(-1, 30)[rows]) + 0.5 * F.bce_loss(other["at_out"], lw)
            loss.backward()
            arena.step()
            return loss
        desc = (f"PMAM post-pre-training (config/pmam/post_pretrain.yaml: PaSST_CNN, LoRA r=8 backbone, CNN branch, TransformerXL d=384, block "
                f"masking + prototype BCE on the masked frames + AT loss), batch={B}/GPU of 10 s @ 32 kHz clips, fwd + bwd + all-reduce + fused AdamW "
                "on the trainable (LoRA / CNN / decoder / head) parameters")
        return net, ext, arena, train_step, desc, wav_host
    if name == "dasm":
        from transformer4sed_b200.src_models.detect_any_sound.detect_any_sound import DASM
        K = 407
        kw = dict(cnn_param=dict(n_in_channel=1, activation="cg", conv_dropout=0.0, kernel_size=[3] * 10, padding=[1] * 10, stride=[1] * 10,
                                 nb_filters=list(schema.PMAM_FILTERS), pooling=[list(p) for p in schema.PMAM_POOLING]),
                  backbone_param=dict(embed_dim=768, passt_feature_layer=10, pretrain_model_path=None,
                                      lora_config=dict(r=8, lora_alpha=1, requires_grad_pretrain=False)),
                  at_param=dict(at_decoder_layer=2, query_projector=True, query_dim=768, out_type="sigmoid", query=None), mlm_dict=None,
                  backbone_upsample_ratio=10, decoder_dim=384, num_heads=12, decoder="transformerXL", decoder_layer_num=3, decoder_pos_emd_len=1000,
                  decoder_expand_rate=1, class_num=K)
        net = DASM(**copy.deepcopy(kw))
        net.load_state_dict(synth.synth_state_dict_like(net, 1))
        net = net.to(dev).train()
        ext = net.get_feature_extractor().eval()
        arena = ParamArena(net, generic(net), shadow_bf16=shadow)
        g = torch.Generator().manual_seed(7 + rank)
        query = (torch.nn.functional.normalize(torch.randn(K, 768, generator=g), dim=-1) * 3).to(dev)
        labels = (torch.rand(B, K, 1000, generator=g) < 0.05).float().to(dev)
        lw = labels.amax(-1)

        def train_step(wav):
            mel = ext.logmel(wav)
            s_, w_, o_ = net(mel, temp_w=4.0, query=query)
            loss = F.bce_loss(s_, labels) + 0.5 * F.bce_loss(w_, lw) + 0.5 * F.bce_loss(o_["at_out"], lw)
            loss.backward()
            arena.step()
            return loss
        desc = (f"DASM open-vocabulary path (K={K} text/audio queries, LoRA backbone, CNN branch, 2-layer tagging decoder, TransformerXL d=384, "
                f"query x frame scores), batch={B}/GPU of 10 s @ 32 kHz clips, fwd + BCE strong/weak/AT + bwd + all-reduce + fused AdamW")
        return net, ext, arena, train_step, desc, wav_host
    raise SystemExit(f"unknown workload {name}")


def run_ours(args):
    import torch.distributed as dist
    from transformer4sed_b200 import _lib, functional as F, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the communicator comes up; rank 0's stdout carries ONE JSON line, so the
        # C-level stdout is pointed at stderr while the process group and its first collective are created
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    F.set_precision(args.precision)
    B = args.batch if args.batch else WORKLOAD_BATCH[args.workload]

    torch.manual_seed(1234)
    net, ext, arena, train_step, wl_desc, wav_host = build_workload(args.workload, args.precision, B, dev, rank)
    wav_pinned = [wav_host.clone().pin_memory() for _ in range(2)]
    wav_dev = wav_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        train_step(wav_dev)
    barrier()
    lib = _lib.load()
    # ---- timed region 1: inputs resident in HBM ------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.t4s_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = train_step(wav_dev)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.t4s_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    # host side of one step: wall time until every launch of a step has been enqueued on an idle stream (no synchronisation inside;
    # one step is fewer launches than the driver's queue holds).  The step is GPU-bound as long as this stays below ms_per_step.
    barrier()
    t_h0 = time.perf_counter()
    train_step(wav_dev)
    host_enqueue_ms = (time.perf_counter() - t_h0) * 1e3
    barrier()
    # ---- timed region 2: end to end from pinned host memory (H2D of the clips + D2H of the loss every step) -------
    copy_stream = torch.cuda.Stream(dev)
    bufs = [torch.empty_like(wav_dev) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            bufs[i % 2].copy_(wav_pinned[i % 2], non_blocking=True)
            ready[i % 2].record(copy_stream)

    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    prefetch(0)
    loss_val = 0.0
    for i in range(args.steps):
        torch.cuda.current_stream().wait_event(ready[i % 2])
        if i + 1 < args.steps:
            copy_stream.wait_stream(torch.cuda.current_stream())  # buffer (i+1)%2 was read by step i-1, already enqueued
            prefetch(i + 1)
        loss_val = train_step(bufs[i % 2]).item()                 # D2H read of the step's loss
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    # ---- instrumented step: per-launch device time of the dominant kernel (tcgen05 GEMM) and of the front end ------
    roof, mel_roof, breakdown = None, None, None
    if rank == 0:
        roof, mel_roof, breakdown = instrumented_step(train_step, wav_dev, ext, ops, B)
    elif world > 1:
        train_step(wav_dev)   # the instrumented step contains the gradient all-reduce: every rank has to take part in it
        torch.cuda.synchronize()
    # ---- replicas in sync? every rank started from the same weights and applied the same all-reduced gradients -------------------
    in_sync = None
    if world > 1:
        cs = arena.flat.double().abs().sum().reshape(1)
        ref = cs.clone()
        dist.broadcast(ref, 0)
        diff = (cs - ref).abs() / ref.abs().clamp_min(1e-30)
        dist.all_reduce(diff, op=dist.ReduceOp.MAX)
        in_sync = {"max_rel_checksum_diff_vs_rank0": float(diff.item()), "ok": bool(diff.item() == 0.0)}
    cpu_base = strict = eager = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_reference_sample(target_seconds=15.0)
    if rank == 0 and world == 1 and args.workload == "matsed" and not args.no_extra_legs:
        del train_step, net, arena
        torch.cuda.empty_cache()
        strict = strict_mode_throughput(B, dev, rank)
        eager = gpu_eager_baseline(dev)
    if rank == 0:
        clips = B * world * args.steps
        pk, pk_src = peaks()
        value = clips / (ms / 1e3)
        names = dict(matsed="10s-clips/sec fwd+bwd MAT-SED", matsed_finetune2="10s-clips/sec MAT-SED mean-teacher step (finetune2)",
                     pmam="10s-clips/sec fwd+bwd PMAM post-pre-training", dasm="10s-clips/sec fwd+bwd DASM (K=407)")
        cfg = {"workload": wl_desc, "workload_name": args.workload, "global_batch": B * world, "per_gpu_batch": B,
               "sample_rate_note": "BASELINE.json says 16 kHz; every reference recipe is 32 kHz and the model asserts 1000 frames (SURVEY §0.1, §8d)",
               "parallelism": f"dp{world}", "l2_policy": "inputs larger than L2 (82 MB of clips + GBs of activations per step)",
               "loss": float(loss_val), "peaks": pk_src}
        if args.workload == "matsed":
            cfg["tc_frac_of_sustained_bf16"] = value / world * 3 * FWD_GFLOP_PER_CLIP / 1e3 / pk["bf16_tflops_sustained"]
        elif roof is not None:
            cfg["tc_frac_of_sustained_bf16_gemm_flops_only"] = roof["gemm_tflop_per_step"] / (ms / args.steps / 1e3) / pk["bf16_tflops_sustained"]
        cfg["host_enqueue_ms_per_step"] = host_enqueue_ms
        if in_sync is not None:
            cfg["params_in_sync"] = in_sync
        if strict is not None:
            cfg["strict_mode_clips_per_s"] = strict
        out = {
            "metric": names[args.workload], "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16": "bf16", "tf32": "tf32", "tf32x3": "tf32x3"}[args.precision], "data": "synthetic",
            "config": cfg,
            "clocks": clocks,
            "e2e": {"value": clips / (ms_e2e / 1e3), "unit": "clips/s", "h2d_bytes_per_step": B * N_SAMPLES * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": roof, "roofline_mel": mel_roof, "kernel_time_breakdown_ms": breakdown,
            "cpu_baseline": cpu_base, "gpu_eager_baseline": eager,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def strict_mode_throughput(B, dev, rank):
    """What the <=1e-3 / argmax-exact contract costs: the same MAT-SED step in the tf32 and tf32x3 precision modes (3 timed steps each)."""
    from transformer4sed_b200 import functional as F
    res = {}
    for mode in ("tf32", "tf32x3"):
        try:
            F.set_precision(mode)
            net, ext, arena, step, _, wav_host = build_workload("matsed", mode, B, dev, rank)
            wav = wav_host.to(dev)
            for _ in range(2):
                step(wav)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                step(wav)
            e1.record()
            torch.cuda.synchronize()
            res[mode] = B * 3 / (e0.elapsed_time(e1) / 1e3)
        except Exception as e:  # noqa: BLE001
            res[mode] = f"failed: {type(e).__name__}: {e}"[:200]
        finally:
            F.set_precision("bf16")
            net = ext = arena = step = None
            torch.cuda.empty_cache()
    res["note"] = "same model / batch / step as the bf16 line; the parity contract (<=1e-3 rel, exact argmax) is asserted in tf32x3 (tests/test_model_gpu.py)"
    return res


def gpu_eager_baseline(dev):
    """BASELINE.md §4.5 / SURVEY §8(d): the reference's own PyTorch modules on this GPU.  The reference does not travel to the GPU box, so
    the oracle port of it (plain torch ops: oracle/model.py + oracle/frontend.py, the code the parity tests pin to the reference) runs the
    same model / loss / fwd+bwd in eager fp32 and under autocast(bfloat16), at the largest batch of {32, 16, 8, 4} that fits.  A reported
    baseline leg outside every timed region; nothing of the product path goes through it."""
    try:
        from oracle import frontend as OF
        from oracle import model as OM
        from transformer4sed_b200 import schema
        from transformer4sed_b200.utils import synth
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    out = {}
    sd = {k: v.to(dev).requires_grad_(True) for k, v in synth.synth_state_dict(schema.mat_sed_shapes(), 4).items()}
    for tag, ctx in (("fp32", None), ("bf16_autocast", torch.bfloat16)):
        for batch in (32, 16, 8, 4):
            try:
                wav = synth.synth_wav(4, N_SAMPLES, seed=1234).repeat(batch // 4, 1).to(dev)
                y, yw = make_labels(batch, 99)
                y, yw = y.to(dev), yw.to(dev)

                def step():
                    for v in sd.values():
                        v.grad = None
                    with torch.autocast("cuda", dtype=ctx, enabled=ctx is not None):
                        mel = OF.passt_logmel(wav)
                        strong, weak, other = OM.mat_sed_forward(mel, sd, decoder_layers=3)
                    loss = OM.bce(strong.float(), y) + 0.5 * OM.bce(weak.float(), yw) + 2.0 * OM.bce(other["at_out"].float(), yw)
                    loss.backward()
                for _ in range(2):
                    step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    step()
                e1.record()
                torch.cuda.synchronize()
                out[tag] = {"clips_per_s": batch * 3 / (e0.elapsed_time(e1) / 1e3), "batch": batch}
                break
            except torch.cuda.OutOfMemoryError:
                torch.cuda.empty_cache()
                continue
            except Exception as e:  # noqa: BLE001
                out[tag] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
                break
        torch.cuda.empty_cache()
    out["what"] = "oracle port of the reference modules (plain PyTorch eager) on this B200: fwd + same losses + bwd, no optimizer step"
    return out


def instrumented_step(train_step, wav_dev, ext, ops, B):
    """One extra step with every C-ABI launch bracketed by CUDA events on the launching stream (never the timed region)."""
    from transformer4sed_b200 import _lib
    pk, pk_src = peaks()
    prof = _lib.LaunchProfiler()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _lib.profiler = prof
    try:
        torch.cuda.synchronize()
        t0.record()
        train_step(wav_dev)
        t1.record()
        torch.cuda.synchronize()
    finally:
        _lib.profiler = None
    step_ms = t0.elapsed_time(t1)
    gemm_ms = gemm_flops = 0.0
    n_gemm = 0
    attn_ms = attn_flops = 0.0
    attn_parts = {}
    # algorithmic flops of the fused attention calls (head size 64): forward 4 N^2 hd per (clip, head) (S = Q K^T and O = P V), backward
    # 2.5 x (five products); Transformer-XL: + the position scores 2 N (2N-1) hd forward, backward 2.7 x forward
    fl = {"t4s_attn_fwd": lambda B_, H, N: 4.0 * N * N * 64 * H * B_, "t4s_attn_bwd": lambda B_, H, N: 10.0 * N * N * 64 * H * B_,
          "t4s_relattn_fwd": lambda B_, H, N: (4.0 * N * N + 2.0 * N * (2 * N - 1)) * 64 * H * B_,
          "t4s_relattn_bwd": lambda B_, H, N: 2.7 * (4.0 * N * N + 2.0 * N * (2 * N - 1)) * 64 * H * B_}
    for name, key, s, e in prof.records:
        if name == "t4s_gemm":
            M, N, K, nb = key[:4]
            gemm_ms += s.elapsed_time(e)
            gemm_flops += 2.0 * M * N * K * nb
            n_gemm += 1
        elif name in fl:
            dt = s.elapsed_time(e)
            attn_ms += dt
            f = fl[name](*key) if key is not None else 0.0
            attn_flops += f
            a = attn_parts.setdefault(name, [0, 0.0, 0.0])
            a[0] += 1
            a[1] += dt
            a[2] += f
    summary = prof.summary()
    shapes = {}
    for name, key, s, e in prof.records:
        if name == "t4s_gemm":
            M, N, K, nb = key[:4]
            a = shapes.setdefault(str(key), [0, 0.0, 0.0])
            a[0] += 1
            a[1] += s.elapsed_time(e)
            a[2] += 2.0 * M * N * K * nb
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        with open(os.environ.get("T4S_BREAKDOWN", "gpurun_out/step_breakdown.json"), "w") as f:
            json.dump({"step_ms": step_ms, "ops": {k: {"launches": v[0], "ms": round(v[1], 4)} for k, v in summary.items()},
                       "gemm_tflops": {k: {"launches": v[0], "ms": round(v[1], 4), "tflops": round(v[2] / v[1] / 1e9, 1)}
                                       for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][1])}}, f, indent=1)
    except OSError:
        pass
    # front end alone, L2 flushed between iterations
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=wav_dev.device)
    ts = []
    for _ in range(5):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ext.logmel(wav_dev)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    mel_ms = statistics.median(ts)
    mel_bytes = B * (4 * N_SAMPLES + 4 * 128 * 1000)
    roof = {"bound": "tensor", "kernel": "t4s::gemm::gemm_kernel (tcgen05, all GEMM launches of one step)", "achieved": gemm_flops / gemm_ms / 1e9,
            "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": gemm_flops / gemm_ms / 1e9 / pk["bf16_tflops_sustained"],
            # dram__bytes_read.sum + dram__bytes_write.sum of ONE captured launch (fc1 shape M=76160 N=3072 K=768, bias epilogue) from the
            # committed `ncu --set full` capture; algorithmic bytes of that launch: 589.6 MB (profiles/r1j_ncu_full_summary.md)
            "traffic": 537.2e6, "traffic_note": "bytes per launch of the fc1-shape GEMM (359.4 GFLOP); profiles/r1j_ncu_full_summary.md",
            "launches": n_gemm, "gemm_ms_per_step": gemm_ms, "gemm_tflop_per_step": gemm_flops / 1e12, "share_of_step": gemm_ms / step_ms, "peak_source": pk_src + " (sustained bf16)",
            "how": "CUDA events around every t4s_gemm launch of one extra instrumented step; flops = 2MNK per launch"}
    mel_roof = {"bound": "hbm", "kernel": "t4s::mel::mel_kernel (+peak kernel)", "achieved": mel_bytes / mel_ms / 1e6, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": mel_bytes / mel_ms / 1e6 / pk["hbm_gbs"],
                "traffic": 93.75e6 * B / 64, "traffic_note": "ncu --set full, 64 clips per launch: 82.0 MB read + 11.8 MB written (profiles/r1j_ncu_full_summary.md)",
                "ms": mel_ms, "clips_per_s": B / mel_ms * 1e3,
                "algorithmic_bytes_per_clip": 4 * N_SAMPLES + 4 * 128 * 1000, "peak_source": pk_src}
    if attn_ms > 0:
        roof["attention"] = {
            "bound": "tensor", "kernel": "t4s::attn fused attention kernels (tcgen05; head size 64), all launches of one step (incl. delta / dq-finish passes)",
            "achieved": attn_flops / attn_ms / 1e9, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
            "frac": attn_flops / attn_ms / 1e9 / pk["bf16_tflops_sustained"], "ms_per_step": attn_ms,
            "per_call": {k: {"launches": v[0], "ms": round(v[1], 3), "tflops": round(v[2] / v[1] / 1e9, 1)} for k, v in attn_parts.items()},
            "note": "head size 64 is bound by the special-function unit and the shared-memory operand pipe before the tensor pipe: 16 exp2 / clk / SM = "
                    "1024 clk per 128 x 128 tile against 512 clk of MMA, so 0.5 of the tensor roofline is the ceiling of the forward without "
                    "polynomial exponentials (profiles/r2_attention_analysis.md)"}
    return roof, mel_roof, {"step_instrumented": step_ms, "gemm": gemm_ms, "fused_attention": attn_ms, "front_end": mel_ms,
                            "other": step_ms - gemm_ms - attn_ms - mel_ms}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port) on the host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_step_fn(batch, threads):
    from oracle import frontend as OF
    from oracle import model as OM
    from transformer4sed_b200 import schema
    from transformer4sed_b200.utils import synth
    torch.set_num_threads(threads)
    sd = synth.synth_state_dict(schema.mat_sed_shapes(), 4)
    for v in sd.values():
        v.requires_grad_(True)
    wav = synth.synth_wav(min(batch, 4), N_SAMPLES, seed=1234).repeat((batch + 3) // 4, 1)[:batch]
    y, yw = make_labels(batch, 99)

    def step():
        for v in sd.values():
            v.grad = None
        mel = OF.passt_logmel(wav)
        strong, weak, other = OM.mat_sed_forward(mel, sd, decoder_layers=3)
        loss = OM.bce(strong, y) + 0.5 * OM.bce(weak, yw) + 2.0 * OM.bce(other["at_out"], yw)
        loss.backward()
        return loss.item()

    return step


def cpu_reference_sample(target_seconds=15.0):
    threads = os.cpu_count() or 1
    step = cpu_step_fn(1, threads)
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter() - t0
    batch = max(1, min(8, int(target_seconds / max(t1, 1e-3))))
    if batch > 1:
        step = cpu_step_fn(batch, threads)
        t0 = time.perf_counter()
        step()
        t1 = time.perf_counter() - t0
    return {"value": batch / t1, "unit": "clips/s", "cores": threads, "kind": "port",
            "sample": f"1 fwd+bwd step of the fp32 CPU oracle (oracle/model.py, same model/loss) on {batch} clip(s), {t1:.1f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    probe = cpu_step_fn(1, threads)
    t0 = time.perf_counter()
    probe()
    t1 = time.perf_counter() - t0
    budget = 150.0
    batch = max(1, min(4, int(budget / ((args.steps + args.warmup) * max(t1, 1e-3)))))
    step = cpu_step_fn(batch, threads)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = step()
    dt = time.perf_counter() - t0
    value = batch * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": "10s-clips/sec fwd+bwd MAT-SED", "value": value, "unit": "clips/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"MAT-SED base, same model / losses as the CUDA arm, bounded sample: {batch} clip(s) of 10 s @ 32 kHz per step on the "
                               "host CPU (the reference is pure PyTorch; timed through the oracle port because /root/reference does not travel)",
                   "global_batch": batch, "loss": loss},
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} timed steps x {batch} clip(s), torch.set_num_threads({threads})"},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="clips per GPU per step (default: 64 MAT-SED, 32 PMAM, 8 DASM)")
    ap.add_argument("--workload", default="matsed", choices=["matsed", "matsed_finetune2", "pmam", "dasm"])
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the strict-mode and GPU-eager baseline legs (rank 0, N=1, after the timed regions)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32", "tf32x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
