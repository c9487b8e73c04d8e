#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native VyomAI hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo (sm_100a kernels)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

Metric (BASELINE.json): caption-train samples/s — one step = one full training step (ViT encoder +
RoPE/GQA decoder forward, shifted cross-entropy, backward, gradient all-reduce, global-norm clip,
AdamW) of the image-text fusion captioner `VisionLanguageModel` on one batch of synthetic,
right-padded image/caption pairs, bf16 weights with fp32 master copies. Data-parallel: every rank
processes its own batch of the same size (weak scaling); the only collective is the NCCL gradient
all-reduce. One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: NCCL's own version / debug prints go to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

WORKLOAD = "captioner_train_vit224p16L4_dec768L8gqa4_rope_rpad_S128"
PER_GPU_BATCH = 64
TEXT_LEN = 127  # + 1 image token = 128 decoder positions
CPU_SAMPLE_BATCH = 8


class TextCfg:
    hidden_size = 768
    num_attention_heads = 12
    num_key_value_heads = 4
    max_position_embeddings = 514
    num_hidden_layers = 8
    vocab_size = 50265
    hidden_dropout_prob = 0.0  # fused path implements p = 0; the reference default is 0.1
    initializer_range = 0.02
    intermediate_size = 3072
    layer_norm_eps = 1e-05
    hidden_act = "gelu"
    pad_token_id = 1


class VitCfg:
    hidden_size = 768
    num_attention_heads = 12
    image_size = (224, 224)
    patch_size = (16, 16)
    num_channels = 3
    num_hidden_layers = 4
    hidden_dropout_prob = 0.0
    initializer_range = 0.02
    intermediate_size = 3072
    layer_norm_eps = 1e-05
    hidden_act = "gelu"


def synth_batch(batch, seed, pin):
    g = torch.Generator().manual_seed(seed)
    pixels = torch.rand((batch, 3, 224, 224), generator=g)
    ids = torch.randint(3, TextCfg.vocab_size, (batch, TEXT_LEN), generator=g)
    lens = torch.randint(8, TEXT_LEN + 1, (batch,), generator=g)
    mask = (torch.arange(TEXT_LEN)[None, :] < lens[:, None]).long()
    ids = torch.where(mask.bool(), ids, torch.full_like(ids, TextCfg.pad_token_id))
    if pin:
        pixels, ids, mask = pixels.pin_memory(), ids.pin_memory(), mask.pin_memory()
    return pixels, ids, mask


# ---- second workload: the notebook-II form of the captioner (Examples/vyom-ai-accelerate-multimodel-2t4.ipynb cell 1) ----
# ViT-base (12 layers) -> all 197 tokens scattered into the <image> slots of a 248-token sequence ([bos, <Caption>] + 197 x
# <image> + caption + eos, right-padded), 8-layer RoPE decoder with the DeBERTa-v3-base widths the notebook takes its config
# from (768 / 12 heads MHA / 3072, vocab 128100 + <image>, LayerNorm eps 1e-7), causal x padding mask, AdamW + clip 1.0.
SLOTS_WORKLOAD = "captioner_II_train_vit224p16L12_197img_in_248tok_dec768L8mha_rope_rpad"
SLOTS_SEQ, SLOTS_IMG, SLOTS_BATCH, SLOTS_IMAGE_TOKEN = 248, 197, 32, 128001


class SlotTextCfg(TextCfg):
    num_key_value_heads = None
    vocab_size = 128101
    layer_norm_eps = 1e-07
    pad_token_id = 0


class SlotVitCfg(VitCfg):
    num_hidden_layers = 12


def synth_slot_batch(batch, seed, pin):
    g = torch.Generator().manual_seed(seed)
    pixels = torch.rand((batch, 3, 224, 224), generator=g)
    n_text = SLOTS_SEQ - 2 - SLOTS_IMG  # 49 caption positions incl. eos
    ids = torch.full((batch, SLOTS_SEQ), SlotTextCfg.pad_token_id, dtype=torch.long)
    ids[:, 0], ids[:, 1] = 1, 5
    ids[:, 2:2 + SLOTS_IMG] = SLOTS_IMAGE_TOKEN
    text = torch.randint(6, 128000, (batch, n_text), generator=g)
    lens = torch.randint(8, n_text + 1, (batch,), generator=g)
    keep = torch.arange(n_text)[None, :] < lens[:, None]
    ids[:, 2 + SLOTS_IMG:] = torch.where(keep, text, torch.full_like(text, SlotTextCfg.pad_token_id))
    mask = (ids != SlotTextCfg.pad_token_id).long()
    if pin:
        pixels, ids, mask = pixels.pin_memory(), ids.pin_memory(), mask.pin_memory()
    return pixels, ids, mask


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            }
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.1)
        except Exception as e:  # NVML missing: report that instead of inventing clocks
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d["bf16_tflops_sustained"], "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's CPU path
# ------------------------------------------------------------------------------------------------
def oracle_state_dict(TextCfg=TextCfg, VitCfg=VitCfg, head="decoder.lm_head.", head_proj="decoder"):
    """Random-init weights of the bench's captioner for the CPU arm, keyed like the reference's state_dict (the names the
    oracle reads), built with plain torch — none of this repo's modules are involved in the reference arm."""
    g = torch.Generator().manual_seed(0)
    H, V, FF = TextCfg.hidden_size, TextCfg.vocab_size, 4 * TextCfg.hidden_size
    kv = (TextCfg.num_key_value_heads or TextCfg.num_attention_heads) * (H // TextCfg.num_attention_heads)
    sd = {}

    def lin(name, out_f, in_f):
        bound = in_f ** -0.5  # nn.Linear's default init range
        sd[name + ".weight"] = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
        sd[name + ".bias"] = (torch.rand(out_f, generator=g) * 2 - 1) * bound

    def ln(name, n):
        sd[name + ".weight"], sd[name + ".bias"] = torch.ones(n), torch.zeros(n)

    def block(pre, fused_qkv, kv_out):
        if fused_qkv:
            lin(pre + "attention.qkv", 3 * H, H)
        else:
            lin(pre + "attention.query", H, H)
            lin(pre + "attention.key", kv_out, H)
            lin(pre + "attention.value", kv_out, H)
        lin(pre + "attention.out.dense", H, H)
        ln(pre + "attention.out.layernorm", H)
        lin(pre + "feed_forward.intermediate", FF, H)
        lin(pre + "feed_forward.out", H, FF)
        ln(pre + "feed_forward.layernorm", H)

    pdim = VitCfg.num_channels * VitCfg.patch_size[0] * VitCfg.patch_size[1]
    npatch = (VitCfg.image_size[0] // VitCfg.patch_size[0]) * (VitCfg.image_size[1] // VitCfg.patch_size[1])
    sd["encoder.pixel_seq.weight"] = (torch.rand(H, VitCfg.num_channels, *VitCfg.patch_size, generator=g) * 2 - 1) * pdim ** -0.5
    sd["encoder.pixel_seq.bias"] = (torch.rand(H, generator=g) * 2 - 1) * pdim ** -0.5
    sd["encoder.cls_token"] = torch.randn(1, 1, pdim, generator=g)
    sd["encoder.position_embeddings.pos_embeddings"] = torch.randn(1, npatch + 1, pdim, generator=g)
    for i in range(VitCfg.num_hidden_layers):
        block(f"encoder.all_layer.{i}.", True, H)
    sd["decoder.word_embeddings.weight"] = torch.randn(V, H, generator=g)
    for i in range(TextCfg.num_hidden_layers):
        block(f"decoder.all_layer.{i}.", False, kv)
    lin(head + "dense", H, H)
    ln(head + "layer_norm", H)
    sd[head + head_proj + ".weight"] = (torch.rand(V, H, generator=g) * 2 - 1) * H ** -0.5
    sd[head + "bias"] = torch.zeros(V)
    sd = {k: v.requires_grad_(True) for k, v in sd.items()}
    sd[head + head_proj + ".bias"] = sd[head + "bias"]
    return sd


def oracle_train(steps, warmup, batch):
    """Times the reference's algorithm (oracle/vyom_oracle.py restatement, fp32, all host threads) on a
    bounded sample of the same workload: forward + shifted CE + backward + AdamW, `batch` samples/step."""
    from oracle import vyom_oracle as O
    sd = oracle_state_dict()
    params = [v for k, v in sd.items() if v.requires_grad and k != "decoder.lm_head.decoder.bias"]
    opt = torch.optim.AdamW(params, lr=1e-5)
    cfg = O.Cfg(768, 12, 4, 514, 8, 50265, 1e-5, "gelu")
    vcfg = O.Cfg(768, 12, None, 514, 4, 0, 1e-5, "gelu", (224, 224), (16, 16), 3)
    times = []
    for it in range(warmup + steps):
        px, ids, mask = synth_batch(batch, 1000 + it, False)
        labels = ids.masked_fill(mask == 0, -100)
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        logits = O.vlm_forward(sd, cfg, vcfg, px, ids, mask, "rope", "gqa")
        loss = O.cross_entropy_shifted(logits[:, 1:], labels)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return batch / (sum(times) / len(times)), sum(times) / len(times)


def oracle_train_slots(steps, warmup, batch):
    """The same for the notebook-II form: oracle ViT (all 197 tokens) -> slot_vlm_forward (training mask) -> slot_loss -> AdamW."""
    from oracle import vyom_oracle as O
    sd = oracle_state_dict(SlotTextCfg, SlotVitCfg, head="lm_head.", head_proj="vocab")
    params = [v for k, v in sd.items() if v.requires_grad and k != "lm_head.vocab.bias"]
    opt = torch.optim.AdamW(params, lr=1e-5, weight_decay=0.0)
    cfg = O.Cfg(768, 12, None, 514, 8, SlotTextCfg.vocab_size, 1e-7, "gelu")
    vcfg = O.Cfg(768, 12, None, 514, 12, 0, 1e-5, "gelu", (224, 224), (16, 16), 3)
    times = []
    for it in range(warmup + steps):
        px, ids, mask = synth_slot_batch(batch, 2000 + it, False)
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        feats = O.vit_forward(sd, vcfg, px, pre="encoder.")
        logits = O.slot_vlm_forward(sd, cfg, feats, ids, mask, True, SLOTS_IMAGE_TOKEN)
        loss = O.slot_loss(logits, ids, mask, SlotTextCfg.pad_token_id, SLOTS_IMAGE_TOKEN)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return batch / (sum(times) / len(times)), sum(times) / len(times)


def host_threads():
    """All host cores for the CPU arm. torchrun exports OMP_NUM_THREADS=1 to its workers, which made round 1's N > 1 reference
    lines single-threaded: set the intra-op pool explicitly."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        pass
    torch.set_num_threads(n)
    return torch.get_num_threads()


def run_reference(args):
    """Reference arm: the reference's CPU path (oracle port: the reference has no setup.py / pyproject, nothing installs
    under baseline/_ref) on ALL host threads, same model / sequence length / per-step batch as one GPU of our arm. If the
    first (warm-up) step shows that K + W steps of 64 samples would not end within ~4.5 minutes on this host, the per-step
    sample is cut to the largest multiple of 8 that does, and the line says so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    slots = args.workload == "slots"
    train = oracle_train_slots if slots else oracle_train
    ours_batch = SLOTS_BATCH if slots else PER_GPU_BATCH
    probe_n = 2 if slots else CPU_SAMPLE_BATCH
    batch = ours_batch
    t0 = time.perf_counter()
    train(1, 0, probe_n)  # probe: one small step (also pages the libraries in)
    probe = time.perf_counter() - t0
    est = probe * (batch / probe_n) * (args.steps + args.warmup)
    if est > 270.0:
        batch = max(probe_n, int(batch * 270.0 / est) // probe_n * probe_n)
    v, sec = train(args.steps, args.warmup, batch)
    sample = (f"{batch} samples per step (same model, S={SLOTS_SEQ if slots else TEXT_LEN + 1}, fp32; our arm: {ours_batch} per GPU), {args.steps} timed steps after "
              f"{args.warmup} warm-up, {cores} host threads")
    line = {
        "impl": "reference", "metric": "caption_train_samples_per_s", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": SLOTS_WORKLOAD if slots else WORKLOAD, "per_gpu_batch": batch, "per_step_samples": batch,
                   "seq_len": SLOTS_SEQ if slots else TEXT_LEN + 1,
                   "host": "cpu oracle port of the reference path (rank 0 only)"},
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
FAMILIES = (  # kernel-name fragment -> C-ABI entry point it belongs to (first match wins)
    ("dp_barrier", "vy_dp_barrier (includes the wait for the peers)"), ("dp_reduce", "vy_dp_reduce_shard"), ("dp_adamw", "vy_dp_adamw_shard"),
    ("gemm_kernel", "vy_gemm"), ("splitk_reduce", "vy_gemm"), ("attn_fwd_kernel", "vy_attn_fwd"), ("attn_bwd", "vy_attn_bwd"),
    ("attn_dsum", "vy_attn_bwd"), ("add_layernorm_fwd", "vy_add_layernorm_fwd"), ("add_layernorm_bwd", "vy_add_layernorm_bwd"),
    ("norm_bwd", "vy_add_layernorm_bwd"), ("adamw", "vy_adamw"), ("sqnorm", "vy_sqnorm"), ("xent", "vy_softmax_xent"),
    ("colsum", "vy_colsum"), ("embed_bwd", "vy_embed_bwd"), ("embed_fwd", "vy_embed_fwd"), ("patchify", "vy_patchify"),
    ("cast4d", "vy_cast4d"), ("act_bwd", "vy_act_bwd"), ("scale_by_ptr", "vy_scale_by_ptr"), ("attn_decode", "vy_attn_decode"),
    ("decode_step", "vy_decode_step"), ("argmax", "vy_argmax_rows"), ("nccl", "nccl"), ("Memcpy", "memcpy"), ("Memset", "memset"),
)


def family_of(kernel_name):
    for frag, fam in FAMILIES:
        if frag in kernel_name:
            return fam
    return "torch_glue"


def replay_profile(step, steps=2):
    """Per-family device time of the step AS TIMED (CUDA-graph replays under CUPTI through torch.profiler): {family:
    {"calls": per step, "ms": per step}}. Round 1's table came from an eager step with events around every call, whose small
    kernels were inflated and whose sum exceeded ms_per_step."""
    from collections import defaultdict
    step()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
    agg = defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        fam = family_of(ev.name)
        agg[fam][0] += 1
        agg[fam][1] += ev.device_time_total
    return {k: {"calls": n / steps, "ms": t / steps / 1e3} for k, (n, t) in agg.items()}


class DecodeCfg:  # BASELINE config 3: GPT-style CLM, kv-cache, prefill 512 + greedy decode 256, batch 32
    hidden_size = 768
    num_attention_heads = 12
    max_position_embeddings = 1024
    num_hidden_layers = 4
    vocab_size = 50265
    hidden_dropout_prob = 0.0
    layer_norm_eps = 1e-05
    hidden_act = "gelu"
    pad_token_id = 1
    eos_token_id = 2


def decode_config3(dev, attn, batch=32, prefill=512, new_tokens=256):
    """Prefill + greedy decode of BASELINE config 3 through DecoderModel.generate (static cache, one CUDA-graph replay per
    token). Returns prefill tok/s, decode tok/s (device time of the replayed steps), the HBM roofline fraction of a decode
    step (algorithmic bytes = every weight but the embedding table + the kv-cache rows read, at the mean context) and
    generate()'s wall clock."""
    import io
    from contextlib import redirect_stdout
    from vyomai_b200 import DecoderModel
    cfg = type("Cfg3", (DecodeCfg,), {"num_key_value_heads": 4})() if attn == "gqa" else DecodeCfg()
    torch.manual_seed(0)
    with redirect_stdout(io.StringIO()):
        model = DecoderModel(cfg, "rope", "gqa" if attn == "gqa" else None)
    model = model.to(dev).to(torch.bfloat16).eval()
    B, P, N = batch, prefill, new_tokens
    ids = torch.randint(3, cfg.vocab_size, (B, P), device=dev)
    mask = torch.ones((B, P), device=dev, dtype=torch.long)
    hkv = 4 if attn == "gqa" else 12
    model.generate(ids, mask, max_len=N, use_cache=True, use_static_cache=True)  # warm-up: tunes the GEMMs, captures the graph
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = model.generate(ids, mask, max_len=N, use_cache=True, use_static_cache=True)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    kv = model._decode_graph_cache
    with torch.no_grad():
        model(ids, mask, use_cache=True, kv_cache=kv, start_pos=0, _logits_last_only=True)  # (one-time tuning / lazy module loads)
        e0.record()
        model(ids, mask, use_cache=True, kv_cache=kv, start_pos=0, _logits_last_only=True)
        e1.record()
        g = model._decode_graph
        g.pos.fill_(P)
        for _ in range(N - 1):
            g.graph.replay()
        e2.record()
    torch.cuda.synchronize()
    pre_ms, step_us = e0.elapsed_time(e1), e1.elapsed_time(e2) * 1e3 / (N - 1)
    n_params = sum(p.numel() for n, p in model.named_parameters() if "word_embeddings" not in n)
    step_bytes = 2.0 * n_params + 2.0 * B * cfg.num_hidden_layers * 2 * hkv * (P + N / 2) * 64
    hbm = peaks()[0]
    return {"workload": f"decoder_clm_L4_{attn}_B{B}_prefill{P}_decode{N}_bf16_staticcache",
            "prefill_tok_per_s": B * P / (pre_ms / 1e3), "decode_tok_per_s": B / (step_us * 1e-6), "decode_us_per_step": step_us,
            "decode_step_algorithmic_MB": step_bytes / 1e6, "decode_hbm_frac": step_bytes / (step_us * 1e-6) / 1e9 / hbm,
            "generate_tok_per_s_wall": B * N / wall, "tokens_checked": int(out.shape[1]),
            "step_kernel": "vy_decode_step" if getattr(g, "fused", None) is not None else "per-op graph"}

def build_package_model():
    from vyomai_b200 import VisionLanguageModel, Vit
    return VisionLanguageModel(TextCfg(), encoder=Vit(VitCfg()), pos_embedding_type="rope", attention_type="gqa")


def build_slots_model():
    from vyomai_b200 import ImageSlotVisionLanguageModel, Vit
    return ImageSlotVisionLanguageModel(Vit(SlotVitCfg()), SlotTextCfg(), decoder_pos_embedding_type="rope")


def slots_labels(ids, mask):
    from vyomai_b200 import slot_caption_labels
    return slot_caption_labels(ids, mask, SlotTextCfg.pad_token_id, SLOTS_IMAGE_TOKEN)


def workload_of(name):
    from vyomai_b200.trainer import caption_labels
    if name == "slots":
        return {"name": SLOTS_WORKLOAD, "batch": SLOTS_BATCH, "seq": SLOTS_SEQ, "synth": synth_slot_batch, "build": build_slots_model,
                "labels": slots_labels, "wd": 0.0}
    return {"name": WORKLOAD, "batch": PER_GPU_BATCH, "seq": TEXT_LEN + 1, "synth": synth_batch, "build": build_package_model,
            "labels": caption_labels, "wd": 0.01}


def measure_slots(dev, steps, warmup):
    """The notebook-II workload on one GPU, appended to the default line: resident-input training steps (CUDA-graph replays)
    timed with CUDA events, and the oracle port of the same step on the host cores (2 samples per step)."""
    import io
    from contextlib import redirect_stdout
    from vyomai_b200.trainer import Trainer
    wl = workload_of("slots")
    torch.manual_seed(0)
    with redirect_stdout(io.StringIO()):
        model = wl["build"]()
    model = model.to(dev).to(torch.bfloat16).train()
    tr = Trainer(model, lr=1e-5, weight_decay=wl["wd"], max_grad_norm=1.0, use_graph=True)
    px, ids, mask = [t.to(dev) for t in wl["synth"](wl["batch"], 23, False)]
    labels = wl["labels"](ids, mask)
    for _ in range(max(warmup, 3)):
        loss = tr.caption_step(px, ids, mask, labels)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = tr.caption_step(px, ids, mask, labels)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # the notebook's own per-GPU batch (8: `prep(batch_size=8)`), same model and trainer (the step is re-captured for the new shape)
    px8, ids8, mask8 = [t.to(dev) for t in wl["synth"](8, 29, False)]
    labels8 = wl["labels"](ids8, mask8)
    for _ in range(3):
        tr.caption_step(px8, ids8, mask8, labels8)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        tr.caption_step(px8, ids8, mask8, labels8)
    e1.record()
    torch.cuda.synchronize()
    ms8 = e0.elapsed_time(e1) / steps
    cores = host_threads()
    v, sec = oracle_train_slots(1, 1, 2)
    n_params = sum(p.numel() for p in model.parameters())
    return {"workload": wl["name"], "per_gpu_batch": wl["batch"], "seq_len": wl["seq"], "image_tokens": SLOTS_IMG, "ms_per_step": ms,
            "samples_per_s": wl["batch"] / (ms / 1e3), "tokens_per_s": wl["batch"] * wl["seq"] / (ms / 1e3), "params_M": round(n_params / 1e6, 1),
            "final_loss": float(loss), "grad_overwrite": bool(tr.grad_overwrite), "kernels_per_step": tr.graph_kernels,
            "notebook_batch_8": {"ms_per_step": ms8, "samples_per_s": 8 / (ms8 / 1e3)},
            "cpu_port": {"value": v, "unit": "samples/s", "cores": cores, "sample": f"1 timed step of 2 samples after 1 warm-up ({sec:.1f} s/step)"}}


def _timed_graph(fn, iters=20):
    """Device time per call of `fn` replayed from a CUDA graph (3 eager warm-up calls on a side stream first)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fn()
    for _ in range(3):
        graph.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def measure_configs_1_2(dev):
    """BASELINE configs[0] and [1] on this GPU with the oracle port timed beside them on the host cores:
    C1 = EncoderModel (default EncoderConfig, RoPE + GQA kv4), 8 x 128 right-padded tokens, forward + backward of (out . G).sum();
    C2 = Vit 224 / patch 16 (4 layers) + Linear(768, 6) on the CLS row, batch 64, forward. FLOPs: SURVEY.md §8(d)."""
    import io
    from contextlib import redirect_stdout
    from oracle import vyom_oracle as O
    from vyomai_b200 import EncoderConfig, EncoderModel, Vit
    tf_sust = peaks()[2]
    out = {}
    cores = host_threads()
    # ---- C1 ----
    cfg = type("C1", (), dict(vars(EncoderConfig()), num_key_value_heads=4, hidden_dropout_prob=0.0))()
    torch.manual_seed(0)
    with redirect_stdout(io.StringIO()):
        enc = EncoderModel(cfg, pos_embedding_type="rope", attention_type="gqa")
    enc = enc.to(dev).to(torch.bfloat16).train()
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, cfg.vocab_size, (8, 128), generator=g)
    lens = torch.randint(16, 129, (8,), generator=g)
    lens[0] = 128
    mask = (torch.arange(128)[None, :] < lens[:, None]).long()
    cot = torch.randn(8, 128, cfg.hidden_size, generator=g) * mask[..., None]
    ids_d, mask_d, cot_d = ids.to(dev), mask.to(dev), cot.to(dev).to(torch.bfloat16)

    def c1_step():
        for p in enc.parameters():
            p.grad = None
        (enc(ids_d, mask_d).logits * cot_d).sum().backward()

    try:
        ms, how = _timed_graph(c1_step), "cuda-graph replay"
    except Exception as e:  # capture refused: time eager launches instead and say so
        torch.cuda.synchronize()
        for _ in range(3):
            c1_step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            c1_step()
        e1.record()
        torch.cuda.synchronize()
        ms, how = e0.elapsed_time(e1) / 20, f"eager launches (graph capture failed: {type(e).__name__})"
    sd = {k: v.detach().float().cpu().requires_grad_(v.dtype.is_floating_point) for k, v in enc.state_dict().items()}
    ocfg = O.Cfg(cfg.hidden_size, cfg.num_attention_heads, 4, cfg.max_position_embeddings, cfg.num_hidden_layers, cfg.vocab_size,
                 cfg.layer_norm_eps, cfg.hidden_act)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        for v in sd.values():
            v.grad = None
        (O.encoder_forward(sd, ocfg, ids, mask, "rope", "gqa") * cot).sum().backward()
        ts.append(time.perf_counter() - t0)
    cpu_s = min(ts[1:])
    out["C1_encoder_8x128_fwd_bwd"] = {"ms_per_step": ms, "samples_per_s": 8 / (ms / 1e3), "tflops": 159.0 / ms, "frac_of_sustained_bf16": 159.0 / ms / tf_sust,
                                      "algorithmic_gflop_per_step": 159.0, "timed_as": how, "dtype": "bf16",
                                      "cpu_port": {"samples_per_s": 8 / cpu_s, "cores": cores, "dtype": "f32", "sample": "best of 2 steps after 1 warm-up"}}
    del enc
    # ---- C2 ----
    torch.manual_seed(0)
    with redirect_stdout(io.StringIO()):
        vit = Vit(VitCfg())
    head = torch.nn.Linear(768, 6)
    vit, head = vit.to(dev).to(torch.bfloat16).eval(), head.to(dev).to(torch.bfloat16)
    px = torch.rand(64, 3, 224, 224, generator=g)
    px_d = px.to(dev).to(torch.bfloat16)

    def c2_step():
        with torch.no_grad():
            return head(vit(pixel_values=px_d).logits[:, 0])

    try:
        ms2, how2 = _timed_graph(c2_step), "cuda-graph replay"
    except Exception as e:
        torch.cuda.synchronize()
        for _ in range(3):
            c2_step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            c2_step()
        e1.record()
        torch.cuda.synchronize()
        ms2, how2 = e0.elapsed_time(e1) / 20, f"eager launches (graph capture failed: {type(e).__name__})"
    vsd = {k: v.detach().float().cpu() for k, v in vit.state_dict().items()}
    vcfg = O.Cfg(768, 12, None, 514, VitCfg.num_hidden_layers, 0, 1e-5, "gelu", (224, 224), (16, 16), 3)
    with torch.no_grad():
        O.vit_forward(vsd, vcfg, px[:8])
        t0 = time.perf_counter()
        O.vit_forward(vsd, vcfg, px[:16])
        cpu2 = time.perf_counter() - t0
    out["C2_vit224p16_L4_b64_fwd"] = {"ms_per_step": ms2, "images_per_s": 64 / (ms2 / 1e3), "tflops": 759.0 / ms2, "frac_of_sustained_bf16": 759.0 / ms2 / tf_sust,
                                     "algorithmic_gflop_per_step": 759.0, "timed_as": how2, "dtype": "bf16",
                                     "cpu_port": {"images_per_s": 16 / cpu2, "cores": cores, "dtype": "f32", "sample": "one 16-image forward after an 8-image warm-up"}}
    return out


C5_VISION = dict(hidden_size=1152, intermediate_size=4304, num_hidden_layers=27, num_attention_heads=16, num_channels=3, image_size=224,
                 patch_size=14, layer_norm_eps=1e-6)
C5_TEXT = dict(vocab_size=257216, hidden_size=2048, intermediate_size=16384, num_hidden_layers=18, num_attention_heads=8,
               num_key_value_heads=1, head_dim=256, max_position_embeddings=8192, rms_norm_eps=1e-6, rope_theta=10000.0)


def measure_config5(dev, batches=(1, 32), new_tokens=50, cpu=True):
    """BASELINE configs[4]: PaliGemma-scale scratch model (SigLIP 27 x 1152 / 14-pixel patches + Gemma 18 x 2048, 8 q heads,
    1 kv head of 256, GeGLU 16384, vocab 257216, tied head), random init, bf16: prefill of 256 image + 8 text tokens, then
    `new_tokens` greedy tokens through a 384-slot static cache — the procedure of Examples/paligemma.ipynb cell 30.
    HBM floor per decoded token: every weight once (5.02 GB incl. the tied table as lm_head) + the kv rows."""
    from vyomai_b200.models.paligemma import PaliGemmaConfig, PaliGemmaForConditionalGeneration, StaticCache
    out = {"workload": "paligemma_scale_siglip27x1152_gemma18x2048_mqa256_prefill264_decode%d_bf16_staticcache384" % new_tokens}
    try:
        cfg = PaliGemmaConfig(vision_config=dict(C5_VISION), text_config=dict(C5_TEXT), image_token_index=257152, vocab_size=257216,
                              projection_dim=2048, hidden_size=2048, pad_token_id=0)
        old = torch.get_default_dtype()
        torch.set_default_dtype(torch.bfloat16)
        try:
            with torch.device(dev):
                torch.manual_seed(0)
                model = PaliGemmaForConditionalGeneration(cfg)
        finally:
            torch.set_default_dtype(old)
        model.tie_weights()
        model.eval()
        with torch.no_grad():
            for n, p in model.named_parameters():  # N(0, 0.02) like a scratch init; norms near identity
                if p.dim() >= 2:
                    p.normal_(0.0, 0.02)
        n_params = sum(p.numel() for p in model.parameters())
        lm_params = sum(p.numel() for n, p in model.language_model.named_parameters())
        out["params_B"] = round(n_params / 1e9, 3)
        hbm = peaks()[0]
        g = torch.Generator().manual_seed(5)
        for B in batches:
            ids = torch.cat([torch.full((B, 256), 257152, dtype=torch.long), torch.randint(2, 250000, (B, 8), generator=g)], dim=1).to(dev)
            mask = torch.ones((B, 264), dtype=torch.long, device=dev)
            px = torch.rand((B, 3, 224, 224), generator=g).to(dev).to(torch.bfloat16)

            from vyomai_b200 import ops
            from vyomai_b200.models.paligemma import PaliGemmaDecodeGraph

            def run(n_new):
                cache = StaticCache(cfg.text_config, batch_size=B, device=dev, dtype=torch.bfloat16, max_cache_len=384)
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                ev[0].record()
                o = model(input_ids=ids, pixel_values=px, attention_mask=mask, past_key_values=cache, use_cache=True, logits_last_only=True)
                first = ops.argmax_rows(o.logits[:, -1])
                ev[1].record()
                g_ = PaliGemmaDecodeGraph(model, cache, mask)
                g_.tok.copy_(first)
                g_.pos.fill_(264)
                g_.capture()  # (one eager warm-up step + the capture: outside the timed region)
                toks = torch.empty((B, n_new - 1), dtype=torch.long, device=dev)
                torch.cuda.synchronize()
                ev[2].record()
                g_.run(first, 264, n_new - 1, toks)
                ev[3].record()
                torch.cuda.synchronize()
                return ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3]) / (n_new - 1)

            run(4)  # warm-up: weight packing, GEMM tuning, lazy module loads
            pre_ms, step_ms = run(new_tokens)
            kv_bytes = 2.0 * B * 18 * 2 * 1 * (264 + new_tokens / 2) * 256
            step_bytes = 2.0 * lm_params + kv_bytes  # (tied table counted once: it is read as the lm_head)
            out[f"B{B}"] = {"prefill_ms": pre_ms, "prefill_tok_per_s": B * 264 / (pre_ms / 1e3), "decode_ms_per_step": step_ms,
                            "decode_tok_per_s": B / (step_ms / 1e3), "decode_step_algorithmic_GB": step_bytes / 1e9,
                            "decode_hbm_frac": step_bytes / (step_ms / 1e3) / 1e9 / hbm,
                            "timed_as": "prefill: eager launches; decode: one CUDA-graph replay per token; CUDA events"}
        del model
        torch.cuda.empty_cache()
    except Exception as e:  # the headline line must still be printed
        out["error"] = f"{type(e).__name__}: {str(e)[:300]}"
        return out
    if cpu:
        try:
            import psutil
            if psutil.virtual_memory().available < 48e9:
                out["cpu_port"] = {"skipped": "less than 48 GB of free host memory for the fp32 weights (11.7 GB) and activations"}
                return out
            from oracle import vyom_oracle as O
            cores = host_threads()
            t0 = time.perf_counter()
            sd = {}
            gg = torch.Generator().manual_seed(1)

            def w(name, *shape):
                sd[name] = torch.empty(shape).normal_(0.0, 0.02, generator=gg)

            t = C5_TEXT
            w("language_model.model.embed_tokens.weight", t["vocab_size"], t["hidden_size"])
            sd["language_model.lm_head.weight"] = sd["language_model.model.embed_tokens.weight"]
            for i in range(t["num_hidden_layers"]):
                lp = f"language_model.model.layers.{i}."
                w(lp + "self_attn.q_proj.weight", 2048, 2048); w(lp + "self_attn.k_proj.weight", 256, 2048)
                w(lp + "self_attn.v_proj.weight", 256, 2048); w(lp + "self_attn.o_proj.weight", 2048, 2048)
                w(lp + "mlp.gate_proj.weight", 16384, 2048); w(lp + "mlp.up_proj.weight", 16384, 2048); w(lp + "mlp.down_proj.weight", 2048, 16384)
                sd[lp + "input_layernorm.weight"] = torch.zeros(2048); sd[lp + "post_attention_layernorm.weight"] = torch.zeros(2048)
            sd["language_model.model.norm.weight"] = torch.zeros(2048)
            build_s = time.perf_counter() - t0
            cfgd = {"hidden_size": 2048, "image_token_index": 257152, "text": t, "vision": C5_VISION}
            ids = torch.randint(2, 250000, (1, 264), generator=gg)  # text-only prompt of the same length: the decode steps are what is timed
            mask = torch.ones(1, 264, dtype=torch.long)
            cache = ([torch.zeros(1, 1, 384, 256) for _ in range(18)], [torch.zeros(1, 1, 384, 256) for _ in range(18)])
            with torch.no_grad():
                t0 = time.perf_counter()
                lg = O.paligemma_forward(sd, cfgd, ids, None, mask, cache=cache, seen=0, cache_len=384)
                pre_s = time.perf_counter() - t0
                nxt = lg[:, -1].argmax(-1, keepdim=True)
                ts = []
                for s in range(3):
                    mask = torch.cat([mask, torch.ones(1, 1, dtype=mask.dtype)], -1)
                    t0 = time.perf_counter()
                    lg = O.paligemma_forward(sd, cfgd, nxt, None, mask, cache=cache, seen=264 + s, cache_len=384)
                    ts.append(time.perf_counter() - t0)
                    nxt = lg[:, -1].argmax(-1, keepdim=True)
            out["cpu_port"] = {"decode_tok_per_s": 1.0 / min(ts), "prefill_tok_per_s": 264 / pre_s, "cores": cores, "dtype": "f32", "batch": 1,
                               "sample": f"Gemma decoder only (text prompt of 264 tokens), 3 decode steps, best; weights built in {build_s:.0f} s"}
        except Exception as e:
            out["cpu_port"] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
    return out


def _guarded(fn, *a, **kw):
    """The appended measurements must never cost the headline line: a failure is reported in place of the numbers."""
    try:
        return fn(*a, **kw)
    except Exception as e:  # noqa: BLE001
        torch.cuda.synchronize()
        return {"error": f"{fn.__name__}: {type(e).__name__}: {str(e)[:300]}"}


def run_ours(args):
    import torch.distributed as dist
    from vyomai_b200 import _lib
    from vyomai_b200.trainer import HostPrefetcher, Trainer
    wl = workload_of(args.workload)
    caption_labels = wl["labels"]

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()  # raises if the CUDA library is missing: there is no fallback

    torch.manual_seed(0)
    import io
    from contextlib import redirect_stdout
    with redirect_stdout(io.StringIO()):
        model = wl["build"]()
    model = model.to(dev).to(torch.bfloat16).train()
    trainer = Trainer(model, lr=1e-5, weight_decay=wl["wd"], max_grad_norm=1.0, use_graph=not args.no_graph,
                      grad_overwrite=not args.no_grad_overwrite, overlap=not args.no_overlap, bucket_mb=args.bucket_mb,
                      dp_mode=args.dp_mode)

    B = wl["batch"]
    host = [wl["synth"](B, 17 + 1000 * rank + i, True) for i in range(2)]
    px_d, ids_d, mask_d = [t.to(dev) for t in host[0]]
    labels_d = caption_labels(ids_d, mask_d)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        return trainer.caption_step(px_d, ids_d, mask_d, labels_d)

    # end to end: every step's batch comes from pinned host memory (copied on a side stream one step ahead, like a
    # pin_memory DataLoader) and the step's loss is read back to the host before the next step starts
    feeder = HostPrefetcher(host, dev)

    def step_e2e(i):
        px, ids, mask = feeder.next()  # this step's inputs; issues the H2D copy of the next step's batch
        loss = trainer.caption_step(px, ids, mask, caption_labels(ids, mask))
        return float(loss)  # device -> host read of the step's result

    for _ in range(args.warmup):
        step_resident()
    sync()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = L.vy_launch_count() + trainer.replayed_kernels
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step_resident()
    e1.record()
    sync()
    ms = e0.elapsed_time(e1)
    launches = L.vy_launch_count() + trainer.replayed_kernels - launches0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    value = world * B * args.steps / (ms / 1e3)

    # end to end: pinned host inputs -> H2D every step, loss read back every step
    for i in range(2):
        step_e2e(i)
    sync()
    e0.record()
    for i in range(args.steps):
        last = step_e2e(i)
    e1.record()
    sync()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (float(t[0]) / 1e3)
    h2d = sum(x.numel() * x.element_size() for x in host[0])

    # Algorithmic work per kernel family: one instrumented EAGER step (the C-ABI wrappers count FLOPs / bytes per call; its
    # timings are not used). Time per family: CUPTI over the captured step as it is replayed in the timed region.
    _lib.TIMER = _lib.KernelTimer()
    trainer._caption_body(px_d, ids_d, mask_d, labels_d)
    work = _lib.TIMER.summary()
    _lib.TIMER = None
    prof = replay_profile(step_resident) if not args.no_graph else {k: {"calls": v["calls"], "ms": v["ms"]} for k, v in work.items()}
    hbm, tf_burst, tf_sust, src = peaks()
    traffic = None  # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    for tname in ("r02_gemm_traffic.json", "r01_gemm_traffic.json"):  # r02: mean over two whole steps (350 launches); r01: 6 launches
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch_mean")
            break
    # the device barriers of the data-parallel step only wait for the other ranks (and absorb the profiler's skew): not work
    ours = {k: v for k, v in prof.items() if k.startswith("vy_") and not k.startswith("vy_dp_barrier")}
    total_ms = sum(d["ms"] for k, d in prof.items() if not k.startswith("vy_dp_barrier"))
    name, d = max(ours.items(), key=lambda kv: kv[1]["ms"])
    w = work.get(name, {"flops": 0.0, "bytes": 0.0})
    timed_as = "cuda-graph replay (CUPTI)" if not args.no_graph else "eager step (CUDA events per call)"
    if w["flops"] > 0:
        achieved = w["flops"] / (d["ms"] / 1e3) / 1e12
        roof = {"bound": "tensor", "kernel": name, "achieved": achieved, "peak": tf_sust, "unit": "TFLOP/s",
                "frac": achieved / tf_sust, "traffic": traffic if name == "vy_gemm" else None,
                "peak_source": f"{src} (sustained bf16 GEMM)", "timed_as": timed_as,
                "launches_per_step": d["calls"], "share_of_kernel_time": d["ms"] / total_ms}
    else:
        achieved = w["bytes"] / (d["ms"] / 1e3) / 1e9
        roof = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                "traffic": None, "peak_source": src, "timed_as": timed_as, "launches_per_step": d["calls"],
                "share_of_kernel_time": d["ms"] / total_ms}
    breakdown = {}
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        wk = work.get(k, {"flops": 0.0, "bytes": 0.0})
        breakdown[k] = {"calls": round(v["calls"], 1), "ms": round(v["ms"], 3),
                        "tflops": round(wk["flops"] / (v["ms"] / 1e3) / 1e12, 1) if wk["flops"] else None,
                        "gbs": round(wk["bytes"] / (v["ms"] / 1e3) / 1e9, 1) if wk["bytes"] else None}

    final_loss = float(loss)
    poisoned = _lib.lib().vy_gemm_poisoned()
    if poisoned != 0:
        raise RuntimeError(f"vy_gemm_poisoned() = {poisoned}: a wait inside a GEMM kernel timed out, the numbers above are void")
    decode = None
    if world == 1 and not args.no_decode:
        # the metric also names decode tok/s: BASELINE config 3, GQA and MHA, on this GPU (inference replicas do not interact,
        # so it is measured at N = 1 only)
        decode = {a: _guarded(decode_config3, dev, a) for a in ("gqa", "mha")}
    small = None
    if world == 1 and args.workload == "package" and not args.no_configs_1_2:
        small = _guarded(measure_configs_1_2, dev)  # BASELINE configs[0], [1]
    c5 = None
    if world == 1 and args.workload == "package" and not args.no_config5:
        c5 = measure_config5(dev, cpu=not args.no_cpu_baseline)  # BASELINE configs[4]
    slots = None
    if world == 1 and args.workload == "package" and not args.no_slots:
        slots = _guarded(measure_slots, dev, args.steps, args.warmup)  # BASELINE configs[3] in its notebook-II form, same GPU
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = host_threads()
            nb = CPU_SAMPLE_BATCH if args.workload == "package" else 2
            v, sec = (oracle_train if args.workload == "package" else oracle_train_slots)(1, 1, nb)
            cpu = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                   "sample": f"1 timed step of {nb} samples after 1 warm-up step (oracle port, fp32, {sec:.1f} s/step)"}
        line = {
            "metric": "caption_train_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": wl["name"], "per_gpu_batch": B, "global_batch": B * world, "seq_len": wl["seq"],
                       "parallelism": f"dp{world}", "dp_step": trainer.dp_mode, "l2": "per-step working set (GBs of activations) exceeds the 126 MB L2",
                       "dropout": 0.0, "optimizer": "AdamW fp32 master + clip 1.0",
                       "cuda_graph": not args.no_graph, "grad_overwrite": bool(trainer.grad_overwrite)},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "clocks": sampler.result(), "roofline": roof, "cpu_baseline": cpu,
            "kernel_breakdown_ms": breakdown, "kernel_breakdown_sum_ms": round(total_ms, 3),
            **({"kernel_breakdown_note": "rank 0 under the profiler; the device barriers' time is the wait for the other ranks (profiling skews "
                                         "them) and is left out of kernel_breakdown_sum_ms and of the roofline's share"} if world > 1 else {}),
            "decode": decode, "notebook_II": slots, "configs_1_2": small, "config_5": c5,
            "final_loss": final_loss, "e2e_last_loss": last,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # NCCL communicators that were captured into CUDA graphs do not always tear down cleanly
        # (destroy_process_group was seen to hang after the result line was out): synchronise, then leave.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="package", choices=["package", "slots"],
                    help="package: VisionLanguageModel of the reference package (1 image token + 127 text, the headline since round 1); "
                         "slots: the notebook-II form (197 image tokens in a 248-token sequence)")
    ap.add_argument("--no-configs-1-2", action="store_true", help="skip the encoder (C1) / ViT (C2) measurements appended at N = 1")
    ap.add_argument("--no-config5", action="store_true", help="skip the PaliGemma-scale prefill + decode measurement appended at N = 1")
    ap.add_argument("--no-slots", action="store_true", help="skip the notebook-II measurement appended at N = 1")
    ap.add_argument("--no-decode", action="store_true", help="skip the config-3 decode measurement appended at N = 1")
    ap.add_argument("--no-grad-overwrite", action="store_true", help="zero + accumulate every gradient instead of overwrite mode")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-overlap", action="store_true", help="one gradient all-reduce after backward instead of bucketed overlap")
    ap.add_argument("--dp-mode", default=None, choices=["p2p", "nccl"],
                    help="N > 1: p2p = sharded optimizer step over NVLink peer memory (default), nccl = bucketed all-reduce + full AdamW")
    ap.add_argument("--bucket-mb", type=float, default=64.0, help="gradient bucket size of the overlapped all-reduce")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
