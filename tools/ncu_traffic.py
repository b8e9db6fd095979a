"""DRAM traffic per launch of the hot kernels from an `ncu --set full` report -> JSON keyed by the kernel names bench.py
uses in its per-kernel table (the L(g)-sized launch of each kernel = the one with the largest traffic)."""
import csv
import json
import subprocess
import sys

NAMES = [("gemm_gather_bf16x3_kernel<256, false>|gather", "gemm_gather<256>+gather+stats"),
         ("egc_forward_kernel", "egc_forward<gate_is_m>"),
         ("egc_backward_dst_kernel", "egc_backward_dst (part of egc_backward(dst+src))"),
         ("egc_backward_src_kernel", "egc_backward_src (part of egc_backward(dst+src))"),
         ("wgrad_bf16x3_kernel", "wgrad<256,256>"),
         ("bn_backward_reduce_kernel", "bn_backward_reduce"),
         ("affine_silu_residual_kernel", "affine_silu_residual")]
def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


rep = sys.argv[1]                 # (the first report names the source of the per-conv kernels; more reports may follow)
per = {}
for one in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", one, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    ik = hdr.index("Kernel Name")
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    units = rows[1]
    for r in rows[2:]:
        name = r[ik]
        b = to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw])
        per.setdefault(name, []).append((b, r[it] + " " + units[it]))
res = {}
gemm = sorted(((max(v)[0], k, max(v)[1]) for k, v in per.items() if "gemm_gather" in k), reverse=True)
for k, v in per.items():
    b, t = max(v)
    short = k.split("(")[0].replace("void ", "").replace("alignn::", "")
    res[short] = {"dram_bytes_per_launch": b, "duration_under_ncu": t, "launches_in_report": len(v), "source": rep.split("/")[-1]}
# aliases under the names bench.py's KernelTimer uses
alias = {}
for short, e in res.items():
    if "egc_forward_kernel" in short or "egc_forward_ring_kernel" in short:
        alias["egc_forward<gate_is_m>"] = e
    if "wgrad_bf16x3_kernel<256, 256>" in short:
        alias["wgrad<256,256>"] = e
    if "bn_backward_reduce_kernel<256>" in short:
        alias["bn_backward_reduce"] = e
    if "affine_silu_residual_kernel<256>" in short:
        alias["affine_silu_residual"] = e
    if "wgrad_batch_kernel<256, 256>" in short:
        alias["wgrad_batch<256>"] = e
    if "ln_silu_forward_kernel<256>" in short:
        alias["ln_silu_forward"] = e
    if "ln_silu_backward_kernel<256>" in short:
        alias["ln_silu_backward"] = e
d = [e for s, e in res.items() if "egc_backward_dst" in s]
s_ = [e for s, e in res.items() if "egc_backward_src" in s]
if d and s_:
    alias["egc_backward(dst+src)"] = {"dram_bytes_per_launch": d[0]["dram_bytes_per_launch"] + s_[0]["dram_bytes_per_launch"],
                                       "source": d[0]["source"], "note": "sum of the two kernels of one call"}
# the gather GEMM launches of one conv, in launch order: [P = node GEMM, gate GEMM (gather + stats), ..., node data
# gradient, edge data gradient (+ residual)]: the two L(g)-sized ones are the first and the last above 300 MB
big = [(b, t) for k, v in per.items() if "gemm_gather" in k for (b, t) in v if b > 3e8]
if big:
    src = rep.split("/")[-1]
    alias["gemm_gather<256>+gather+stats"] = {"dram_bytes_per_launch": big[0][0], "duration_under_ncu": big[0][1], "source": src,
                                               "note": "gate GEMM of the L(g) conv forward (first L(g)-sized launch)"}
    alias["gemm_gather<256>+residual"] = {"dram_bytes_per_launch": big[-1][0], "duration_under_ncu": big[-1][1], "source": src,
                                           "note": "edge data-gradient GEMM of the L(g) conv backward (last L(g)-sized launch)"}
res.update(alias)
print(json.dumps(res, indent=1))
