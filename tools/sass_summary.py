"""Static evidence for profiles/: per-kernel registers / stack / spills from the ptxas logs of the in-tree build and
SASS opcode histograms of the hot kernels (cuobjdump -sass of xray_projection_render_b200/lib/libcuda_render.so).
python tools/sass_summary.py > profiles/sass_r2_summary.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
BUILD = ROOT / "xray_projection_render_b200" / "csrc" / "build"
LIB = ROOT / "xray_projection_render_b200" / "lib" / "libcuda_render.so"
HOT = [r"render_span_kernel<\(int\)1, \(bool\)0, \(bool\)0, \(bool\)0, \(int\)5>", r"render_span_kernel<\(int\)1, \(bool\)0, \(bool\)1, \(bool\)1, \(int\)4>",
       r"span_bin_kernel", r"render_async_kernel<\(int\)2, \(int\)1, \(bool\)0, \(int\)5>", r"render_fast_kernel<\(int\)2, \(int\)1, \(bool\)0, \(bool\)0, \(int\)3>",
       r"render_volume_tex_kernel<\(int\)32, \(int\)1, \(int\)0>", r"render_volume_tex_kernel<\(int\)32, \(int\)1, \(int\)1>",
       r"render_volume_f64_kernel<\(int\)0, \(int\)0>", r"render_scene_exact_kernel", r"voxelize_tiles_kernel", r"voxelize_cells_kernel"]


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def ptxas_table():
    rows = []
    for log in sorted(BUILD.glob("*.ptxas.log")):
        cur = None
        stack = spill_s = spill_l = 0
        for line in log.read_text().splitlines():
            m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
            if m:
                cur, stack, spill_s, spill_l = m.group(1), 0, 0, 0
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and cur and stack == spill_s == spill_l == 0:
                stack, spill_s, spill_l = (int(v) for v in m.groups())
                continue
            m = re.search(r"Used (\d+) registers", line)
            if m and cur:
                rows.append((log.name.split(".")[0], cur, int(m.group(1)), stack, spill_s, spill_l))
                cur = None
    return rows


def main():
    rows = ptxas_table()
    dm = demangle([r[1] for r in rows])
    print("== ptxas (nvcc 12.9, -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo): registers / stack bytes / spill stores / spill loads per entry kernel")
    for f, name, regs, stack, ss, sl in rows:
        short = re.sub(r"\(xr::RenderParams.*", "", dm.get(name, name)).replace("void ", "")
        print(f"{f:14s} {regs:4d} regs {stack:5d} B stack {ss:5d} B st {sl:5d} B ld  {short}")
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    names = [f.split("\n", 1)[0].strip() for f in funcs]
    dmf = demangle(names)
    print("\n== SASS opcode histograms of the hot kernels (static instruction counts; cuobjdump -sass)")
    for pat in HOT:
        for n, body in zip(names, funcs):
            d = dmf.get(n, n)
            if not re.search(pat, d):
                continue
            ops = collections.Counter()
            for line in body.splitlines():
                m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(?:\.\S+)?\s", line)
                if m:
                    ops[m.group(1)] += 1
            tot = sum(ops.values())
            short = re.sub(r"\(xr::RenderParams.*", "", d).replace("void ", "")
            top = ", ".join(f"{k} {v}" for k, v in ops.most_common(22))
            fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"))
            fp32 = sum(v for k, v in ops.items() if k in ("FFMA", "FADD", "FMUL", "FSETP", "FMNMX", "FSEL", "FCHK"))
            mem = {k: v for k, v in ops.items() if k in ("LDG", "STG", "LDS", "STS", "LDL", "STL", "LDC", "TLD4", "TEX", "ATOMG", "ATOMS", "RED", "UTMALDG")}
            print(f"\n{short}\n  {tot} instructions; fp64 arithmetic {fp64}, fp32 arithmetic {fp32}, MUFU {ops.get('MUFU', 0)}, memory {mem}\n  {top}")
            break
    tc = [k for k in ("HMMA", "UTCHMMA", "UTCQMMA", "UTMALDG", "TCGEN05") if re.search(r"\b" + k, sass)]
    print("\n== tensor-core / TMA mnemonics present in the product library:", tc or "none (nothing on this path is a contraction; the voxel path reads through TLD4 -- DESIGN.md section 5)")


if __name__ == "__main__":
    sys.exit(main())
