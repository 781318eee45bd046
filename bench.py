#!/usr/bin/env python
"""bench.py -- homomorphic convolutions / second of the evalConv_BN hot interval
(eval.go:250-260 = conv_then_pack + bias add) on B200.

A "step" is one pass of the fused path over one batch of --cts independent level-1 input
ciphertexts (synthetic uniform residues) with the workload of BASELINE.json
configs[1]: conv k=3, B=16 output channels, N=2^16, level 1 -> 0, one special prime.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--cts M] [--batch B]
    python bench.py --impl reference ...      # the CPU path (oracle port) on the host cores

N>1 runs under torchrun, one rank per GPU; ranks process independent ciphertexts (weak
scaling, no data-path collective); the only collectives are the barrier and the max of the
device times.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from optimal_conv_b200 import params as PR  # noqa: E402
from optimal_conv_b200 import shard, synth  # noqa: E402

N = 1 << PR.LOGN
LIMB = N * 8
METRIC = "homomorphic convolutions/sec (N=2^16 CKKS, k=3, batch=16)"


def conv_alg_limbs(B):
    """algorithmic limbs of one conv: 10B - 1 + 4 log2 B"""
    return 10 * B - 1 + 4 * (B.bit_length() - 1)


def kernel_limbs(name, M, B, first_launch_only=False, algorithmic=False, deferred=False):
    """Limbs fused kernel `name` reads + writes, summed over its launches in one run (Stage A: 1 launch; Stage B: one
    per pack level) or for its first launch.
    algorithmic=False: DISTINCT limbs the launch moves, scratch included (DESIGN.md "Kernels"; what ncu's dram__bytes
    should show when nothing is fetched twice).
    algorithmic=True: SURVEY.md 8(d) -- only operands and results of the reference operations the kernel completes
    (input ciphertexts, plaintexts, key slices, output ciphertexts); half-transformed scratch counts for nothing, so
    extra passes over a limb show up as a lower achieved figure.
    deferred=True: the plan carries level-0 polynomials as pairs (U, e): the U half stands for the
    reference's ciphertext in the algorithmic count, the e half is scratch."""
    na, jobs = B, M * B * 2
    if deferred:
        if name == "A1":   # k_convA1 (with at least one pack level the U halves of stage A are never written)
            return 2 * M + na + (0 if algorithmic else jobs)
        if name == "A2":
            return 0 if algorithmic else 2 * jobs                          # w1 -> e
        if name == "F1":
            return 0 if algorithmic else 4 * M
        if name == "F2":
            return 2 * M + 1 + (0 if algorithmic else 4 * M)               # result, bias (<- w, U)
    else:
        if name == "A1":
            return 2 * M + na + (0 if algorithmic else jobs)      # ct limb q1 of both polys, pt limb q1 per channel (-> w1)
        if name == "A2":
            return 0 if algorithmic else 2 * jobs                  # w1 -> w2
        if name == "A3":
            return 2 * M + na + jobs + (0 if algorithmic else jobs)  # ct limb q0, pt limb q0, level-0 output (<- w2)
    tot, n = 0, na
    while n > 1:
        nbt = M * (n // 2)                  # butterflies in this level's launch
        if deferred and algorithmic:
            tot += {"B1": 2 * nbt, "B2": 0, "B3": 2, "B4": 0, "B5": 6 * nbt + 3}[name]  # U halves of a, b (4), out (2); monomial, 2 key Q limbs
        elif deferred and n == na:          # first level: U halves formed from ct_in and per-butterfly plaintext tables (pairs)
            tot += {"B1": M + 2 * (na // 2) + nbt,               # ct_in.c1 limb q0, tables -> w1
                    "B2": 5 * nbt, "B3": 3 * nbt + 2, "B4": 8 * nbt,
                    "B5": 3 * nbt + 2 * M + 4 * (na // 2) + 4}[name]  # w4, ct_in limb q0 x2, two tables, key Q limbs as pairs -> 2 U out
        elif deferred:
            tot += {"B1": 4 * nbt + 2,      # Ua1, Ub1, monomial pairs -> z, w1
                    "B2": 5 * nbt,          # w1, ea1, eb1 -> w2 (p0), w4 (q0)
                    "B3": 3 * nbt + 2,
                    "B4": 8 * nbt,          # per polynomial: w3, ea, eb -> e out
                    "B5": 7 * nbt + 6}[name]  # w4, z, Ua0, Ua1, Ub0 -> 2 U out; monomial pairs, 2 key Q limbs as pairs
        elif algorithmic:
            tot += {"B1": 2 * nbt, "B2": 0, "B3": 2, "B4": 0,            # a1, b1 | - | key P limbs | -
                    "B5": 6 * nbt + 3 + (1 if n == 2 else 0)}[name]      # a, b (4), out (2); monomial, 2 key Q limbs (+ bias)
        else:
            tot += {"B1": 4 * nbt + 2,      # a1, b1, monomial pairs (16 B per coefficient) -> z, w1
                    "B2": 2 * nbt,
                    "B3": 3 * nbt + 2,      # w2 (shared by both key polys), 2 key P limbs -> 2 outputs
                    "B4": 4 * nbt,
                    "B5": 8 * nbt + 4 + (1 if n == 2 else 0)}[name]  # w4 x2, z, a0, a1, b0, 2 out; monomial pairs, 2 key Q limbs (+ bias)
        if first_launch_only:
            return tot
        n //= 2
    return tot


def group_alg_limbs(M, B):
    """SURVEY.md 8(d) bytes of the NTT + key-switch kernel group (B1..B5) on the first pack level: the two input
    ciphertexts and the output of every butterfly, the level's key slice, the monomial"""
    nbt = M * (B // 2)
    return 6 * nbt + 4 + 1


NCU_SUMMARY = {64: "r02c_ncu_summary.csv"}   # capture of the default `python bench.py` command with HEC_DEFER=0
NCU_SUMMARY_DEFERRED = {64: "r02e_ncu_summary.csv"}   # the same with deferred transforms (the default plan)


def kname(short, deferred):
    """kernel symbol of a plan launch (the first, largest, of a run): the deferred plan reuses k_convA1 / k_convB3 and
    has its own A2, B1 (first level), B2, B4, B5, F1, F2"""
    if deferred:
        return {"A1": "k_convA1", "B1": "k_defB1f", "B3": "k_convB3", "B5": "k_defB5<1>"}.get(short, "k_def" + short)
    return "k_conv" + short


def ncu_traffic(kernel, cts, deferred=False):
    """dram bytes (read + write) of the first launch of `kernel` from the committed ncu --set full capture of the
    same command, or None."""
    import csv
    table = NCU_SUMMARY_DEFERRED if deferred else NCU_SUMMARY
    if cts not in table:
        return None
    p = os.path.join(ROOT, "profiles", table[cts])
    try:
        rows = list(csv.reader(open(p)))
        h, units = rows[0], rows[1]                       # ncu scales units per column: the second row names them
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        ir, iw = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
        for r in rows[2:]:
            if r[0].startswith(kernel + "(") or r[0].startswith("void " + kernel + "("):
                return float(r[ir]) * mult[units[ir]] + float(r[iw]) * mult[units[iw]]
    except Exception:
        pass
    return None


# measured integer ceiling of the path (profiles/r01_ubench_int_pipe.txt): Shoup modmul/clk/SM
INT_MODMUL_PER_CLK_SM = 3.91


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi -lms sampling of SM clocks and throttle reasons during the timed region
    (B200_PROFILING.md clocks line)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 6:
                self.rows.append(f)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.rows[0][1]),
                "reasons": reasons, "samples": len(self.rows)}


def config4(ctx0, torch, rank, world):
    """BASELINE.json config 4: conv 3, batch 256, the ciphertexts of a step dealt over the GPUs (8 per GPU per step);
    device-resident conv/s of this rank (the caller's line carries the per-N figure; ranks are independent)."""
    from optimal_conv_b200 import hec
    B, M, steps = 256, 8, 4
    Q, P = PR.Q_SET6[:2], PR.P_PACK
    ctx0.close()
    ctx = hec.Context(PR.LOGN, Q, P, device=torch.cuda.current_device())
    w = synth.conv_workload(Q, P, PR.LOGN, B, seed=2025, n_ct=1)
    mono = np.zeros((PR.LOGN, N), dtype=np.uint64)
    for i in range(PR.LOGN):
        m = np.zeros(N, dtype=np.uint64)
        m[1 << i] = 1
        mono[i] = ctx.ntt(m, 0)
    ker = [ctx.upload_pt(w["pt_ker"][i], PR.SCALE) for i in range(B)]
    idx = [ctx.upload_pt(mono[i:i + 1], 1.0) for i in range(PR.LOGN)]
    bias = ctx.upload_pt(w["bias"][None, :], PR.SCALE)
    for j, k in w["keys"].items():
        ctx.upload_swk((1 << (j + 1)) + 1, k, 0)
    cts = [ctx.upload_ct(synth.uniform_limbs(900 + 2 * m + 100 * rank, Q, N), synth.uniform_limbs(901 + 2 * m + 100 * rank, Q, N), PR.SCALE)
           for m in range(M)]
    plan = ctx.plan(ker, 1, PR.SCALE, PR.SCALE, idx, bias, M)
    for _ in range(3):
        plan.run(cts)
    torch.cuda.synchronize()
    ctx.timer_start()
    for _ in range(steps):
        plan.run(cts)
    ms = ctx.timer_stop_ms()
    v, _ = shard.throughput(M, steps, ms, device="cuda")
    plan.destroy()
    ctx.close()
    return {"workload": "conv3_B256_N65536_lvl1to0_P1", "cts_per_step_per_gpu": M, "steps": steps, "value": v, "unit": "conv/s",
            "n_gpus": world, "note": "device resident, whole job over all ranks; inputs of 8 ciphertexts reused (16 MiB, L2-resident; "
                                     "the 5 GiB of intermediates are not)"}


def run_reference(args):
    """The reference's CPU implementation of the path: the oracle port (the reference is Go +
    an un-vendored module and cannot be built here).  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle.orc import Ct, Oracle
    B = args.batch
    Q, P = PR.Q_SET6[:2], PR.P_PACK
    w = synth.conv_workload(Q, P, PR.LOGN, B, seed=2024, n_ct=1)
    o = Oracle(PR.LOGN, Q, P)
    idx = o.monomial_pts()
    ct = Ct(w["ct"][0][0], w["ct"][0][1], PR.SCALE)
    ncpu = os.cpu_count() or 1

    def one(nt):
        t = time.perf_counter()
        o.conv_then_pack(ct, w["pt_ker"], PR.SCALE, 1, PR.SCALE, idx, w["keys"], w["bias"], nthreads=nt)
        return time.perf_counter() - t

    # the reference is single-threaded; also try all host threads over channels / tree nodes
    t1 = min(one(1) for _ in range(2))
    tall = min(one(ncpu) for _ in range(2)) if ncpu > 1 else t1
    nt = 1 if t1 <= tall else ncpu
    for _ in range(args.warmup):
        one(nt)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one(nt)
    dt = time.perf_counter() - t0
    val = args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "conv/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "conv3_B%d_N65536_lvl1to0_P1" % B, "cts_per_step": 1},
        "cpu_baseline": {"value": val, "unit": "conv/s", "cores": nt, "kind": "port",
                         "sample": "1 conv (B=%d) per step, %d steps; 1-thread %.3f s/conv, %d-thread %.3f s/conv"
                                   % (B, args.steps, t1, ncpu, tall)},
        "e2e": {"value": val, "unit": "conv/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_side_workload(args):
    """Measured numbers for the other rows of SURVEY.md section 8 (op-level, generic kernels):
    --workload conv_bl    : one evalConv_BN_BL_test interval (eval.go:108-131), set 7, level 1, alpha = 2,
                            for --batch B (4,16,64,256) and --ker k (3,5,7): k^2-1 hoisted rotations +
                            B/2 x (k^2 MulNew/Add + RotateNew) + bias  (BASELINE.json config 3: the k sweep)
    --workload keyswitch  : KeySwitcher.SwitchKeysInPlace at level 27, alpha = 5, beta = 6 (SURVEY.md 8d stress)
    --workload mul_relin  : MulRelinNew(ct, ct) + Rescale at level 10, alpha = 5 (the multiply of
                            evalReLU's polynomial evaluation, conv.go:435-480)
    --workload eval_relu  : the whole evalReLU on one level-15 ciphertext (three EvaluatePoly + the final multiply)
    --workload bootstrap_ctos : BootstrappConv_CtoS, the first half of the split bootstrapping, full modulus chain
    --workload prep_ker   : the plaintext loop of prep_Ker (conv.go:510-515): --batch B float coefficient vectors from
                            HOST memory -> EncodeCoeffs + ToNTT on the device -> B level-1 plaintexts"""
    import torch
    from optimal_conv_b200 import hec
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    if args.workload == "resnet20":
        # BASELINE.json configs[4]: the 20-layer chain of `resnet 3 20 1` (test.go:76-366), one hec_conv_bn_relu per layer.
        # Keys, masks and bootstrapper matrices are synthetic with the real supports (optimal_conv_b200/resnet.py): this
        # measures time and memory of a real inference; logits need the Go host's secret key.
        from optimal_conv_b200 import resnet
        free0, total = torch.cuda.mem_get_info()
        net = resnet.Resnet(hec, depth=20, ker_wid=args.ker, log=lambda *a: print(*a, file=sys.stderr))
        free1, _ = torch.cuda.mem_get_info()
        image = np.random.default_rng(5).uniform(0, 1, 32 * 32 * 3)
        recs, peak = [], 0
        for it in range(args.warmup + args.steps):
            ct, rec = net.run(image, seed=it)
            peak = max(peak, total - torch.cuda.mem_get_info()[0])
            g0, g1 = ct.download()
            ct.free()
            if it >= args.warmup:
                recs.append(rec)
        import hashlib
        dig = hashlib.sha256(g0.tobytes() + g1.tobytes()).hexdigest()[:16]
        per = [{"layer": recs[0][i]["layer"], "kind": recs[0][i]["kind"], "conv": recs[0][i]["conv"], "level_out": recs[0][i]["level_out"],
                "eval_ms": statistics.median(r[i]["eval_ms"] for r in recs), "prep_ms": statistics.median(r[i]["prep_ms"] for r in recs)}
               for i in range(len(recs[0]))]
        tot_eval, tot_prep = sum(p["eval_ms"] for p in per), sum(p["prep_ms"] for p in per)
        line = {"metric": "encrypted ResNet-20 inferences/sec (resnet 3 20 1 layer chain, N=2^16, synthetic keys and matrices with the real supports)",
                "value": 1e3 / (tot_eval + tot_prep), "unit": "images/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": tot_eval + tot_prep, "higher_is_better": True, "dtype": "u64", "data": "synthetic",
                "config": {"workload": "resnet20_k%d" % args.ker, "layers": len(per)},
                "eval_ms_total": tot_eval, "host_prep_ms_total": tot_prep, "layers": per,
                "hbm_bytes": {"resident_after_setup": int(free0 - free1), "peak_during_run": int(peak), "device_total": int(total)},
                "setup_s": net.setup_s, "rotation_keys_uploaded": len(net.have), "final_ciphertext_sha256": dig,
                "note": "per layer: host_prep = prep_Ker float reshaping + device EncodeCoeffs/ToNTT; eval = one hec_conv_bn_relu call"}
        net.close()
        print(json.dumps(line))
        return
    if args.workload == "keyswitch":
        Q, P, level = PR.Q_SET6, PR.P_ALL, 27
        ctx = hec.Context(PR.LOGN, Q, P)
        g = ctx.galois_for_rotation(7)
        key = np.stack([np.stack([synth.uniform_limbs(6000 + 3 * d + kk, Q + P, N) for kk in range(2)]) for d in range(6)])
        ctx.upload_swk(g, key, level)
        a0, a1 = synth.uniform_limbs(1, Q, N), synth.uniform_limbs(2, Q, N)
        A = ctx.upload_ct(a0, a1, PR.SCALE)
        out = ctx.CopyNew(A)

        def step():
            ctx.RotateGal(A, g, out)
        unit, name = "rotations/s", "RotateGal (key-switch + automorphism) at level 27, alpha=5, beta=6"
        # L (c1) + 2*beta*(L+alpha) (key) + 2L (out) + L (c0) limbs, SURVEY.md 8d
        alg = (28 + 2 * 6 * 33 + 2 * 28 + 28) * LIMB
    elif args.workload == "prep_ker":
        Q, P, B = PR.Q_SET6[:2], PR.P_PACK, args.batch
        ctx = hec.Context(PR.LOGN, Q, P)
        vals = (synth.splitmix64(99, B * N).astype(np.float64) / 2.0 ** 63 - 1.0).reshape(B, N) / 9.0

        def step():
            for pt in ctx.EncodeCoeffsNTTMany(vals, 1, PR.SCALE):
                pt.free()
        unit, name = "layers/s", "prep_Ker plaintext loop (conv.go:510-515): %d x EncodeCoeffs + ToNTT at level 1, host floats in" % B
        # per plaintext: N doubles read, 2 limbs written by the scaling kernel, 2 limbs through the transform (in + out)
        alg = B * (N * 8 + 2 * LIMB + 2 * 2 * LIMB)
        if args.cpu_sample > 0:
            from oracle.orc import Oracle
            o = Oracle(PR.LOGN, Q, P)
            t0 = time.perf_counter()
            o.encode_coeffs_ntt(vals[0], PR.SCALE, 1)
            cpu_s = (time.perf_counter() - t0) * B
    elif args.workload == "mul_relin":
        level = 10
        Q, P = PR.Q_SET6[:level + 1], PR.P_ALL
        ctx = hec.Context(PR.LOGN, Q, P)
        beta = (level + 1 + len(P) - 1) // len(P)
        rlk = np.stack([np.stack([synth.uniform_limbs(8000 + 10 * d + kk, Q + P, N) for kk in range(2)]) for d in range(beta)])
        ctx.upload_rlk(rlk, level)
        a = [synth.uniform_limbs(61 + t, Q, N) for t in range(4)]
        A, Bc = ctx.upload_ct(a[0], a[1], PR.SCALE), ctx.upload_ct(a[2], a[3], PR.SCALE)

        def step():
            r = ctx.MulRelinNew(A, Bc)
            ctx.Rescale(r, PR.SCALE)
            r.free()
        unit, name = "multiplications/s", "MulRelinNew(ct, ct) + Rescale at level %d, alpha=5, beta=%d" % (level, beta)
        L_ = level + 1
        alg = (4 * L_ + 2 * beta * (L_ + 5) + 2 * L_) * LIMB   # two cts in, relinearisation key, one ct out
        if args.cpu_sample > 0:
            from oracle.orc import Ct, Oracle
            o = Oracle(PR.LOGN, Q, P)
            t0 = time.perf_counter()
            o.rescale(o.mul_relin(Ct(a[0], a[1], PR.SCALE), Ct(a[2], a[3], PR.SCALE), rlk), PR.SCALE)
            cpu_s = time.perf_counter() - t0
    elif args.workload == "bootstrap_ctos":
        Q, P = PR.Q_SET6, PR.P_ALL
        ctx = hec.Context(PR.LOGN, Q, P)
        # --diagonals real: the supports of the bootstrapper's real factor matrices at LogSlots 15 (16 + 31 + 31 + 15
        # diagonals, synth.dft_factor_specs) instead of the 3-5 diagonals per factor of the parity-test operands; --log-slots
        # picks a sparse bootstrapper (14..11, main.go:60-83)
        specs = synth.dft_factor_specs(args.log_slots, 4, 27) if args.diagonals == "real" else None
        if specs is None:
            keys, kconj, rlk, b = synth.ctos_operands(N)
        else:   # one seeded key buffer stands in for every rotation key (timing: sizes and levels are the real ones)
            _, kconj, rlk, b = synth.ctos_operands(N, specs=[])
            rots = set()
            for n1, diags, _ in specs:
                rots |= {d % n1 for d in diags if d % n1} | {(d // n1) * n1 for d in diags if d // n1}
            rots |= {1 << i for i in range(args.log_slots, PR.LOGN - 1)}
            keys = {r: kconj for r in sorted(rots)}
            tile = [(synth.uniform_limbs(7000 + t, Q, N), synth.uniform_limbs(7500 + t, P, N)) for t in range(3)]
            b["mats"] = [({d: (tile[(i + j) % 3][0][:ml + 1], tile[(i + j) % 3][1]) for j, d in enumerate(diags)}, n1, ml, float(Q[ml]))
                         for i, (n1, diags, ml) in enumerate(specs)]
        for r, k in keys.items():
            ctx.upload_swk(ctx.galois_for_rotation(r), k, 27)
        ctx.upload_swk(2 * N - 1, kconj, 27)
        ctx.upload_rlk(rlk, 27)
        mats = [ctx.upload_ptdiag(args.log_slots if specs else PR.LOGN - 1, n1, ml, ms, D) for D, n1, ml, ms in b["mats"]]
        a0, a1 = synth.uniform_limbs(61, Q[:2], N), synth.uniform_limbs(62, Q[:2], N)
        A = ctx.upload_ct(a0, a1, PR.SCALE * 2.0 ** 8)

        def step():
            g0, g1, _ = ctx.BootstrappConv_CtoS(A, b, mats)
            g0.free()
            if g1 is not None:   # sparse packing returns one ciphertext
                g1.free()
        unit = "half-bootstraps/s"
        name = ("BootstrappConv_CtoS (eval.go:447-459) over the 28+5 modulus chain of set 6: modUp, 4 hoisted linear transforms "
                "(%s), degree-63 Chebyshev sine + double angle, alpha=5"
                % ("factor matrices with the real diagonal supports at LogSlots %d: %s diagonals, %d rotation keys"
                   % (args.log_slots, "+".join(str(len(sp[1])) for sp in specs), len(keys)) if specs else "synthetic sparse factors, 3-5 diagonals each"))
        alg = None
        if args.cpu_sample > 0 and specs is None:
            from oracle.orc import Ct, Oracle
            o = Oracle(PR.LOGN, Q, P)
            t0 = time.perf_counter()
            o.bootstrapp_conv_ctos(Ct(a0, a1, PR.SCALE * 2.0 ** 8), b, keys, kconj, rlk)
            cpu_s = time.perf_counter() - t0
    elif args.workload == "eval_relu":
        level = 15
        Q, P = PR.Q_SET6[:level + 1], PR.P_ALL
        ctx = hec.Context(PR.LOGN, Q, P)
        beta = (level + 1 + len(P) - 1) // len(P)
        rlk = np.stack([np.stack([synth.uniform_limbs(8000 + 10 * d + kk, Q + P, N) for kk in range(2)]) for d in range(beta)])
        ctx.upload_rlk(rlk, level)
        a0, a1 = synth.uniform_limbs(61, Q, N), synth.uniform_limbs(62, Q, N)
        A = ctx.upload_ct(a0, a1, PR.SCALE)

        M_ = max(1, min(args.cts, 16))
        As = [A] + [ctx.upload_ct(synth.uniform_limbs(71 + 2 * t, Q, N), synth.uniform_limbs(72 + 2 * t, Q, N), PR.SCALE) for t in range(M_ - 1)]

        def step():
            for r in ctx.evalReLUMany(As, 0.0, PR.SCALE):
                r.free()
        unit = "ReLU evaluations/s"
        name = "evalReLU (conv.go:435-480) on %d level-15 ciphertexts per call, set 6 ReLU primes, alpha=5: 3 EvaluatePoly (deg 7, 7, 13) + final multiply" % M_
        units_per_step = M_
        alg = None
        if args.cpu_sample > 0:
            from oracle.orc import Ct, Oracle
            o = Oracle(PR.LOGN, Q, P)
            t0 = time.perf_counter()
            o.eval_relu(Ct(a0, a1, PR.SCALE), 0.0, rlk, PR.SCALE)
            cpu_s = time.perf_counter() - t0
    else:
        # one evalConv_BN_BL_test interval for the (B, w) row of main.go:578-579 and kernel width --ker
        Q, P = PR.Q_SET7[:2], PR.P_PACK_BL
        ctx = hec.Context(PR.LOGN, Q, P)
        B, k = args.batch, args.ker
        w_ = PR.WIDTHS[PR.BATCHES.index(B)]
        max_batch, rot_step, h = N // (2 * w_ * w_), w_ * w_, k // 2
        rots = [i * w_ + j for i in range(-h, h + 1) for j in range(-h, h + 1)] + [t * rot_step for t in range(1, max_batch)]
        keys = {}
        for r in rots:
            if r and r not in keys:
                keys[r] = np.stack([np.stack([synth.uniform_limbs(5000 + 13 * (r % 9973) + kk, Q + P, N) for kk in range(2)])])
                ctx.upload_swk(ctx.galois_for_rotation(r), keys[r], 1)
        a0, a1 = synth.uniform_limbs(61, Q, N), synth.uniform_limbs(62, Q, N)
        A = ctx.upload_ct(a0, a1, PR.SCALE)
        # a small pool of distinct plaintexts reused cyclically keeps host generation short
        pool_np = [synth.uniform_limbs(700 + t, Q, N) for t in range(16)]
        pool = [ctx.upload_pt(p_, PR.SCALE) for p_ in pool_np]
        taps = [[pool[(i * k * k + t) % 16] for t in range(k * k)] for i in range(max_batch)]
        bias_np = synth.uniform_limbs(99, Q, N)
        bias = ctx.upload_pt(bias_np, PR.SCALE * PR.SCALE)

        def step():
            ctx.conv_bl(A, w_, k, rot_step, taps, bias).free()
        unit = "baseline conv calls/s"
        name = "evalConv_BN_BL_test interval (eval.go:108-131), B=%d (w=%d), k=%d, level 1, alpha=2: %d hoisted + %d full rotations, %d MulNew" % (
            B, w_, k, k * k - 1, max_batch - 1, max_batch * k * k)
        # algorithmic limbs (SURVEY.md 8d per-op figures; L = 2 limbs, alpha = 2, beta = 1): the input ciphertext;
        # per hoisted rotation its key slice 2*beta*(L+alpha) and its output 2L; per tap the rotated ciphertext 2L and
        # the plaintext L; per output channel group the partial sum 2L written, then (but for the first) a full
        # rotation of it: c1 L + key + out 2L + c0 L; the bias L and the result 2L
        L_, ks_ = 2, 2 * 1 * (2 + 2)
        limbs = 2 * L_ + (k * k - 1) * (ks_ + 2 * L_) + max_batch * k * k * 3 * L_ + max_batch * 2 * L_ \
            + (max_batch - 1) * (L_ + ks_ + 2 * L_ + L_) + L_ + 2 * L_
        alg = limbs * LIMB
        if args.cpu_sample > 0:
            from oracle.orc import Ct, Oracle
            o = Oracle(PR.LOGN, Q, P)
            taps_np = [[pool_np[(i * k * k + t) % 16] for t in range(k * k)] for i in range(max_batch)]
            t0 = time.perf_counter()
            o.conv_bl(Ct(a0, a1, PR.SCALE), w_, k, rot_step, taps_np, PR.SCALE, keys, bias_np)
            cpu_s = time.perf_counter() - t0
    for _ in range(max(3, args.warmup)):
        step()
    ctx.sync()
    l0 = ctx.launch_count()
    ctx.timer_start()
    for _ in range(args.steps):
        step()
    ms = ctx.timer_stop_ms()
    line = {"metric": name, "value": args.steps * (units_per_step if args.workload == "eval_relu" else 1) / (ms / 1e3), "unit": unit, "n_gpus": 1, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "dtype": "u64",
            "data": "synthetic", "config": {"workload": args.workload, "path": "op-level generic kernels"},
            "gpu_launches": ctx.launch_count() - l0}
    if args.workload in ("conv_bl", "mul_relin", "eval_relu", "bootstrap_ctos", "prep_ker") and args.cpu_sample > 0 \
            and not (args.workload == "bootstrap_ctos" and args.diagonals == "real"):
        line["cpu_baseline"] = {"value": 1.0 / cpu_s, "unit": unit, "cores": 1, "kind": "port",
                                "sample": "1 call of the same workload, oracle port, 1 thread"}
    if alg:
        peak, src = peaks()
        line["roofline"] = {"bound": "hbm", "achieved": alg / (ms / args.steps / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                            "frac": alg / (ms / args.steps / 1e3) / 1e9 / peak, "traffic": None, "peak_source": src}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cts", type=int, default=64, help="independent input ciphertexts per step per GPU")
    ap.add_argument("--batch", type=int, default=16, help="B: output channels packed per ciphertext")
    ap.add_argument("--ker", type=int, default=3, help="kernel width k (only changes the work of --workload conv_bl)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="conv", choices=["conv", "conv_bl", "keyswitch", "mul_relin", "eval_relu", "bootstrap_ctos", "prep_ker", "resnet20"],
                    help="conv = the headline fused path; the others are op-level side measurements")
    ap.add_argument("--cpu-sample", type=int, default=12, help="convs timed for cpu_baseline (0 = skip)")
    ap.add_argument("--ring", type=int, default=8, help="distinct input batches rotated through (> L2)")
    ap.add_argument("--diagonals", default="few", choices=["few", "real"], help="bootstrap_ctos: factor-matrix supports")
    ap.add_argument("--log-slots", type=int, default=15, help="bootstrap_ctos --diagonals real: LogSlots of the bootstrapper (15 full, 14..11 sparse)")
    ap.add_argument("--check", type=int, default=64, help="ciphertexts of one step compared with the oracle before timing (0 = skip)")
    ap.add_argument("--config4", type=int, default=1, help="also measure BASELINE.json config 4 (B = 256, 8 ciphertexts per GPU per step)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    if args.workload != "conv":
        return run_side_workload(args)

    import torch
    import torch.distributed as dist
    from optimal_conv_b200 import hec

    rank, local, world = shard.world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    # host side of the end-to-end path: this rank's threads and (first-touch) pinned buffers stay on the cores and the
    # memory of the NUMA node its GPU hangs off, each rank on its own share of them
    placement = shard.pin_rank_to_gpu_node(local, world)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, M = args.batch, args.cts
    Q, P = PR.Q_SET6[:2], PR.P_PACK

    # ---- operands (seeded; every rank its own ciphertexts, shared weights/keys) ----
    w = synth.conv_workload(Q, P, PR.LOGN, B, seed=2024, n_ct=1)
    ctx = hec.Context(PR.LOGN, Q, P, device=local)
    mono = np.zeros((PR.LOGN, N), dtype=np.uint64)
    for i in range(PR.LOGN):  # pl_idx[i] = NTT(X^(2^i)) via the GPU NTT (conv.go:248-253)
        m = np.zeros(N, dtype=np.uint64)
        m[1 << i] = 1
        mono[i] = ctx.ntt(m, 0)
    ker = [ctx.upload_pt(w["pt_ker"][i], PR.SCALE) for i in range(B)]
    idx = [ctx.upload_pt(mono[i:i + 1], 1.0) for i in range(PR.LOGN)]
    bias = ctx.upload_pt(w["bias"][None, :], PR.SCALE)
    for j, k in w["keys"].items():
        ctx.upload_swk((1 << (j + 1)) + 1, k, 0)
    # ring of distinct input batches, larger than L2 in total: ring * M * 2 MiB
    ring = max(1, args.ring)
    host_in0 = torch.empty((ring, M, 2, N), dtype=torch.int64).pin_memory()
    host_in1 = torch.empty((ring, M, 2, N), dtype=torch.int64).pin_memory()
    v0, v1 = host_in0.numpy().view(np.uint64), host_in1.numpy().view(np.uint64)
    for r in range(ring):
        for m in range(M):
            s = 7 + 1000 * rank + 100 * r + m
            v0[r, m] = synth.uniform_limbs(2 * s, Q, N)
            v1[r, m] = synth.uniform_limbs(2 * s + 1, Q, N)
    dev_in = [[ctx.upload_ct(v0[r, m], v1[r, m], PR.SCALE) for m in range(M)] for r in range(ring)]
    host_out0 = torch.empty((2, M, N), dtype=torch.int64).pin_memory()  # double buffered: two batches in flight
    host_out1 = torch.empty((2, M, N), dtype=torch.int64).pin_memory()
    plan = ctx.plan(ker, 1, PR.SCALE, PR.SCALE, idx, bias, M)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """sum of per-step device times (CUDA events on the library's stream); L2 flushed between steps"""
        tot = 0.0
        for s in range(steps):
            flush.zero_()
            torch.cuda.synchronize()
            ctx.timer_start()
            fn(s)
            tot += ctx.timer_stop_ms()
        return tot

    def step_dev(s):
        plan.run(dev_in[s % ring])

    def step_e2e_sync(s):
        r = s % ring
        plan.run_host(host_in0[r], host_in1[r], host_out0[0], host_out1[0])

    def e2e_pipelined(steps):
        """K steps through hec_plan_submit_host / hec_plan_wait: every step's inputs cross PCIe (H2D) and its
        outputs come back (D2H) inside the span; copies overlap the neighbouring steps' kernels."""
        plan.span_begin()
        t = -1
        for s in range(steps):
            r, o = s % ring, s & 1
            t = plan.submit_host(host_in0[r], host_in1[r], host_out0[o], host_out1[o])
            if s >= 1:
                plan.wait(t - 1)
        plan.wait(t)
        return plan.span_end_ms()

    # ---- parity of THIS shape before anything is timed: one whole step (M ciphertexts, B channels) downloaded and
    #      compared bit for bit with the CPU oracle (the checker; not inside any timed span) ----
    parity = None
    if args.check > 0:
        import hashlib
        from oracle.orc import Ct, Oracle
        o = Oracle(PR.LOGN, Q, P)
        oidx = o.monomial_pts()
        outs = plan.run(dev_in[0])
        n_chk = min(M, args.check if rank == 0 else 2)
        hsh = hashlib.sha256()
        for m in range(n_chk):
            g0, g1 = outs[m].download()
            ref = o.conv_then_pack(Ct(v0[0, m], v1[0, m], PR.SCALE), w["pt_ker"], PR.SCALE, 1, PR.SCALE, oidx, w["keys"],
                                   w["bias"], nthreads=max(1, (os.cpu_count() or 1) // max(1, world)))[0]
            if not (np.array_equal(g0, ref.c0) and np.array_equal(g1, ref.c1)):
                raise SystemExit("bench.py: GPU result of ciphertext %d differs from the oracle at the bench shape" % m)
            hsh.update(g0.tobytes()); hsh.update(g1.tobytes())
        parity = {"checked_ciphertexts": n_chk, "of": M, "bit_exact_vs_oracle": True, "sha256": hsh.hexdigest()[:16]}
    # ---- kernel-resident throughput ----
    timed(step_dev, args.warmup)
    sampler = ClockSampler(local)
    barrier()
    l0 = ctx.launch_count()
    ms_dev = timed(step_dev, args.steps)
    launches = ctx.launch_count() - l0
    barrier()
    # ---- end to end through the C ABI with host buffers ----
    e2e_pipelined(args.warmup)
    barrier()
    ms_e2e = e2e_pipelined(args.steps)
    barrier()
    ms_e2e_sync = timed(step_e2e_sync, max(3, args.steps // 4)) / max(3, args.steps // 4)
    sampler.stop()
    # ---- latency of a single conv (batch of one ciphertext), device resident ----
    plan1 = ctx.plan(ker, 1, PR.SCALE, PR.SCALE, idx, bias, 1)
    lat = []
    for i in range(13):
        torch.cuda.synchronize()
        ctx.timer_start()
        plan1.run(dev_in[i % ring][:1])
        lat.append(ctx.timer_stop_ms())
    latency_ms = statistics.median(lat[3:])
    plan1.destroy()
    # ---- the per-convolution entry a Go caller binds: hec_conv_then_pack on host buffers -- upload
    #      of the level-1 ciphertext, the call (plan found in the context's cache after the first), download ----
    lat_call = []
    for i in range(13):
        t0 = time.perf_counter()
        ct1 = ctx.upload_ct(v0[i % ring, 0], v1[i % ring, 0], PR.SCALE)
        r1 = ctx.conv_then_pack(ct1, ker, 1, PR.SCALE, idx, bias)
        r1.download()
        lat_call.append(1e3 * (time.perf_counter() - t0))
        ct1.free(); r1.free()
    latency_call_ms = statistics.median(lat_call[3:])
    # ---- per-kernel times (kernels launched one by one) for the roofline of the dominant kernel ----
    prof = {}
    for rep in range(3):
        flush.zero_()
        torch.cuda.synchronize()
        for name, ms in plan.profile(dev_in[rep % ring]):
            prof.setdefault(name, []).append(ms)
    # whole-job throughput: every rank's convs over the slowest rank's device time (no data-path collective)
    value, ms_dev = shard.throughput(M, args.steps, ms_dev, device="cuda")
    e2e, ms_e2e = shard.throughput(M, args.steps, ms_e2e, device="cuda")
    # BASELINE.json config 4 on every rank (its throughput is a collective over the ranks), before the others leave
    cfg4 = config4(ctx, torch, rank, world) if (args.config4 and B == 16) else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    # launches of each kernel per run and jobs per launch
    per_kernel = {}
    for name, times in prof.items():
        reps = 3
        n_launch = len(times) // reps
        per_kernel[name] = {"ms_per_run": sum(times) / reps, "launches_per_run": n_launch}
    # "the dominant kernel": the longest-running one among the kernels that read or write ciphertexts of the reference
    # operation they complete, i.e. whose algorithmic bytes in the sense of SURVEY.md 8(d) grow with the batch (A1, A3, B1,
    # B5); A2, B2, B4 only move half-transformed scratch (0 such bytes) and B3 adds two key limbs to that
    deferred = plan.deferred
    dom = max((k for k in per_kernel if k in ("A1", "A3", "B1", "B5")), key=lambda k: per_kernel[k]["ms_per_run"])
    dom_ms = per_kernel[dom]["ms_per_run"]
    n_l = per_kernel[dom]["launches_per_run"]
    share = dom_ms / sum(v["ms_per_run"] for v in per_kernel.values())
    # roofline of the dominant kernel, per launch, on its first (largest) launch of a run
    first_ms = sum(prof[dom][i * n_l] for i in range(3)) / 3
    alg_bytes_launch = kernel_limbs(dom, M, B, first_launch_only=True, algorithmic=True, deferred=deferred) * LIMB
    distinct_bytes_launch = kernel_limbs(dom, M, B, first_launch_only=True, deferred=deferred) * LIMB
    achieved = alg_bytes_launch / (first_ms / 1e3) / 1e9
    # the NTT + key-switch kernel group of the first pack level (B1..B5 together)
    grp_ms = sum(sum(prof[k][i * per_kernel[k]["launches_per_run"]] for i in range(3)) / 3 for k in ("B1", "B2", "B3", "B4", "B5") if k in prof)
    grp_bytes = group_alg_limbs(M, B) * LIMB
    # integer ceiling: modular multiplies of one conv / measured modmul throughput
    if deferred:   # 1 transform per channel and polynomial, 5 per butterfly, 2 at the end; fewer point-wise products too
        units = 2 * B * 2 + 10 * (B - 1) + 4
        modmuls = units * (N // 2) * 8 + (3 * B * 2 + 8 * (B - 1)) * N
    else:
        units = 4 * B * 2 + 12 * (B - 1)                 # half-transforms (8 stages) per conv
        modmuls = units * (N // 2) * 8 + (5 * B * 2 + 10 * (B - 1)) * N
    sm_mhz = (sampler.summary().get("sm_mhz") or 1965.0)
    int_peak = INT_MODMUL_PER_CLK_SM * 148 * sm_mhz * 1e6
    conv_bytes = conv_alg_limbs(B) * LIMB
    conv_gbs = conv_bytes * (M * args.steps) / (ms_dev / 1e3) / 1e9  # per GPU

    line = {
        "metric": METRIC, "value": value, "unit": "conv/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "conv3_B%d_N65536_lvl1to0_P1" % B, "cts_per_step_per_gpu": M,
                   "convs_per_step": world * M, "l2": "L2 flushed (256 MiB write) between steps + ring of %d input "
                   "batches (%d MiB)" % (ring, ring * M * 4 * LIMB >> 20), "timing": "sum of per-step CUDA-event times, max over ranks"},
        "e2e": {"value": e2e, "unit": "conv/s", "h2d_bytes_per_step": M * 4 * LIMB, "d2h_bytes_per_step": M * 2 * LIMB,
                "ms_per_step": ms_e2e / args.steps, "ms_per_step_unpipelined": ms_e2e_sync,
                "note": "hec_plan_submit_host/hec_plan_wait: every step H2D-copies its pinned level-1 ciphertexts and "
                        "D2H-copies its level-0 results inside the timed span (two steps in flight, inputs fresh from "
                        "the host so no L2 flush applies); kernel plaintexts/keys resident like a reused prep_Ker result"},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "roofline": {"bound": "hbm", "kernel": kname(dom, deferred), "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic(kname(dom, deferred), M, deferred) if B == 16 else None,
                     "peak_source": peak_src, "alg_bytes_per_launch": alg_bytes_launch,
                     "alg_bytes_definition": "SURVEY.md 8(d): operands and results of the reference operations only; scratch excluded",
                     "distinct_bytes_per_launch": distinct_bytes_launch, "launch_ms": first_ms,
                     "launch": "first (largest) of %d launches per run" % n_l, "share_of_step": share,
                     "kernel_selection": "longest among the kernels whose algorithmic bytes grow with the batch (A1, A3, B1, B5); the "
                                         "scratch-to-scratch passes A2, B2, B3, B4 are covered by roofline_group / roofline_conv",
                     "traffic_source": "profiles/%s (ncu --set full, same command)" % (NCU_SUMMARY_DEFERRED if deferred else NCU_SUMMARY).get(M, "-"),
                     "plan": "deferred forward transforms (DESIGN.md 4.5)" if deferred else "transform per rescale / mod-down"},
        "roofline_group": {"kernels": "k_convB1..B5 (NTT + key-switch group), first pack level", "bound": "hbm",
                           "alg_bytes": grp_bytes, "ms": grp_ms, "achieved": (grp_bytes / (grp_ms / 1e3) / 1e9) if grp_ms else None,
                           "peak": peak, "unit": "GB/s", "frac": (grp_bytes / (grp_ms / 1e3) / 1e9 / peak) if grp_ms else None},
        "roofline_int": {"bound": "integer multiply pipe (fmaheavy)", "modmuls_per_conv": modmuls,
                         "achieved": modmuls * (M * args.steps) / (ms_dev / 1e3), "peak": int_peak,
                         "unit": "modmul/s", "frac": modmuls * (M * args.steps) / (ms_dev / 1e3) / int_peak,
                         "peak_source": "tools/ubench_modmul.cu: 3.91 Shoup modmul/clk/SM (profiles/r01_ubench_int_pipe.txt)"},
        "roofline_conv": {"alg_bytes_per_conv": conv_bytes, "achieved": conv_gbs, "peak": peak, "unit": "GB/s",
                          "frac": conv_gbs / peak, "note": "whole conv, per GPU: (10B-1+4log2B) limbs x convs / time"},
        "kernels_ms_per_run": {k: round(v["ms_per_run"], 4) for k, v in sorted(per_kernel.items())},
        "latency_ms_single_conv": latency_ms,
        "latency_ms_single_call": latency_call_ms,
        "latency_note": "single_conv: a prepared plan of one ciphertext, device resident; single_call: hec_ct_upload + "
                        "hec_conv_then_pack (plan cached in the context) + hec_ct_download on pageable host arrays, wall clock",
        "parity": parity,
        "host_placement": placement,
    }
    if cfg4 is not None:
        line["config4"] = cfg4
    # ---- CPU baseline beside it: the oracle port, 1 thread (the reference is single-threaded) ----
    if world == 1 and args.cpu_sample > 0:
        from oracle.orc import Ct, Oracle
        o = Oracle(PR.LOGN, Q, P)
        oidx = o.monomial_pts()
        ct = Ct(w["ct"][0][0], w["ct"][0][1], PR.SCALE)
        o.conv_then_pack(ct, w["pt_ker"], PR.SCALE, 1, PR.SCALE, oidx, w["keys"], w["bias"])
        t0 = time.perf_counter()
        for _ in range(args.cpu_sample):
            o.conv_then_pack(ct, w["pt_ker"], PR.SCALE, 1, PR.SCALE, oidx, w["keys"], w["bias"])
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": args.cpu_sample / dt, "unit": "conv/s", "cores": 1, "kind": "port",
                                "sample": "%d convs of the same workload (B=%d), oracle port, 1 thread of %d host cores"
                                          % (args.cpu_sample, B, os.cpu_count() or 1)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
