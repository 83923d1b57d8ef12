"""Join an ncu launch list (gpu__time_duration per launch) of one eager UNet forward with the engine's recorded
program, so every launch gets its op, shape, algorithmic flops/bytes.  Runs on CPU (plan_only engine)."""
import csv, sys, re, json, collections
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from tiny_config import FULL_UNET
from tooncrafter_b200 import modules
from tooncrafter_b200.engine import UNetEngine

KERNELS_PER_OP = {"conv_gemm": ["tc_gemm_kernel"],  # (all template instantiations share the name)
                  "groupnorm": ["gn_stats_kernel", "gn_apply_kernel"],
                  "layernorm": ["layernorm_kernel"], "row_stats": ["row_stats_kernel"], "attention": ["tc_attn_kernel"],
                  "temporal_attention": ["temporal_attn_mma_kernel"], "time_embed": ["sincos_kernel", "small_linear_kernel", "small_linear_kernel"],
                  "small_linear": ["small_linear_kernel"], "ncthw_to_cl": ["ncthw_to_cl_kernel"], "cl_to_ncthw": ["cl_to_ncthw_kernel"],
                  "upsample2x": ["upsample2x_kernel"], "phase_split2": ["phase_split2_kernel"], "copy2d": ["copy2d_kernel"]}

def describe(fn, a, kw):
    n = fn.__name__
    if n == "conv_gemm":
        a_dims, taps, out_dims, n_cols = a[1], a[4], a[6], a[7]
        M = out_dims[0] * out_dims[1] * out_dims[2]; K = len(taps) * a_dims[3]
        return dict(op="gemm", M=M, N=n_cols, K=K, taps=len(taps), flops=2.0 * M * n_cols * K,
                    bytes=2.0 * (a_dims[0]*a_dims[1]*a_dims[2]*a_dims[3] + n_cols * K + M * (n_cols // 2 if kw.get("geglu") else n_cols)))
    if n == "groupnorm":
        e = kw["frames"] * kw["hw"] * kw["C"]
        return dict(op="groupnorm", elems=e, fps=kw["frames_per_stat"], C=kw["C"], bytes=4.0 * e, flops=0)
    if n in ("layernorm", "row_stats"):
        e = kw["rows"] * kw["C"]; return dict(op="layernorm", elems=e, C=kw["C"], bytes=(4.0 if n == "layernorm" else 2.0) * e, flops=0)
    if n == "attention":
        fl = sum(4.0 * kw["q_batches"] * kw["heads"] * kw["Lq"] * s["Lk"] * 64 for s in a[1])
        return dict(op="attention", B=kw["q_batches"], Lq=kw["Lq"], Lk=[s["Lk"] for s in a[1]], heads=kw["heads"], flops=fl, bytes=0)
    if n == "temporal_attention":
        e = kw["B"] * kw["T"] * kw["P"] * kw["heads"] * 64
        return dict(op="temporal_attention", elems=e, bytes=8.0 * e, flops=4.0 * kw["B"] * kw["P"] * kw["heads"] * kw["T"] ** 2 * 64)
    return dict(op=n, flops=0, bytes=0)

def main(csv_path, B=2, out=None):
    with torch.device("meta"):
        sk = modules.UNetModel(**FULL_UNET)
    m = sk.to_empty(device="cpu")
    eng = UNetEngine(m, device="cpu", plan_only=True)
    plan = eng.plan_for(B, 16, 40, 64, 77 + 256)
    ops_list = [(fn, a, kw) for fn, a, kw in plan.ctx.calls] + [(fn, a, kw) for fn, a, kw in plan.main.calls]
    lines = [l for l in open(csv_path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    # one row per (launch ID, metric): gpu__time_duration.sum and, optionally, dram__bytes_read/write.sum
    per_id = collections.OrderedDict()
    for r in rows:
        name = r["Kernel Name"]
        if "<unnamed>::" not in name or "at::" in name:
            continue
        e = per_id.setdefault(r["ID"], dict(name=re.search(r"<unnamed>::(\w+)", name).group(1), grid=r["Grid Size"], ms=0.0, dram=0.0))
        v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]; mname = r["Metric Name"]
        if mname.startswith("gpu__time_duration"):
            e["ms"] = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}.get(u, 1e-6)
        elif mname.startswith("dram__bytes"):
            e["dram"] += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    ours = [(e["name"], e["ms"], e["grid"], e["dram"]) for e in per_id.values()]
    i = 0; table = []                              # first eager pass: ctx program then main program
    for fn, a, kw in ops_list:
        ks = KERNELS_PER_OP[fn.__name__]
        if fn.__name__ == "groupnorm" and ours[i][0] == "gn_fused_kernel":
            ks = ["gn_fused_kernel"]               # single-pass GroupNorm (small / medium tensors)
        ms = 0.0; dram = 0.0
        for k in ks:
            ok = (ours[i][0] == k or (k == "row_stats_kernel" and ours[i][0] == "layernorm_kernel")   # older captures
                  or (k == "tc_attn_kernel" and ours[i][0] in ("tc_attn3_kernel", "tc_attn_xs_kernel")))   # round-2 kernels
            assert ok, (i, ours[i], k)
            ms += ours[i][1]; grid = ours[i][2]; dram += ours[i][3]; i += 1
        d = describe(fn, a, kw); d["ms"] = ms; d["grid"] = grid; d["dram"] = dram
        table.append(d)
    tot = sum(d["ms"] for d in table)
    agg = collections.defaultdict(lambda: [0.0, 0, 0.0, 0.0, 0.0])
    for d in table:
        key = d["op"]
        if d["op"] == "gemm":
            key = f"gemm M={d['M']} N={d['N']} K={d['K']}"
        elif d["op"] == "groupnorm":
            key = f"groupnorm C={d['C']} elems={d['elems']} fps={d['fps']}"
        elif d["op"] == "attention":
            key = f"attention Lq={d['Lq']} Lk={d['Lk']} heads={d['heads']}"
        elif d["op"] in ("layernorm",):
            key = f"layernorm C={d['C']} elems={d['elems']}"
        elif d["op"] == "temporal_attention":
            key = f"temporal_attention elems={d['elems']}"
        agg[key][0] += d["ms"]; agg[key][1] += 1; agg[key][2] += d["flops"]; agg[key][3] += d["bytes"]; agg[key][4] += d["dram"]
    print(f"total {tot:.3f} ms over {len(table)} ops")
    lines_out = []
    for k, (ms, n, fl, by, dr) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        s = f"{ms:8.3f} ms {100*ms/tot:5.1f}% n={n:3d} {fl/ms/1e9 if ms else 0:8.1f} TF/s {by/ms/1e6 if ms else 0:8.1f} GB/s alg  dram {dr/1e6:9.1f} MB (alg {by/1e6:9.1f})  {k}"
        print(s); lines_out.append(s)
    gem = [d for d in table if d["op"] == "gemm"]
    if gem and sum(d["dram"] for d in gem) > 0:
        summ = dict(launches=len(gem), ms_total=sum(d["ms"] for d in gem), flops_total=sum(d["flops"] for d in gem),
                    dram_bytes_total=sum(d["dram"] for d in gem), dram_bytes_per_launch=sum(d["dram"] for d in gem) / len(gem),
                    algorithmic_bytes_per_launch=sum(d["bytes"] for d in gem) / len(gem), share_of_forward=sum(d["ms"] for d in gem) / tot,
                    source="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, "
                           "one eager UNet forward (B=2), every tc_gemm_kernel launch")
        print(json.dumps(summ))
        if out:
            Path(out).with_suffix(".gemm_traffic.json").write_text(json.dumps(summ, indent=1) + "\n")
    if out:
        Path(out).write_text(f"total {tot:.3f} ms over {len(table)} ops (ncu cold-cache serialized launch times)\n" + "\n".join(lines_out) + "\n")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2, sys.argv[3] if len(sys.argv) > 3 else None)
