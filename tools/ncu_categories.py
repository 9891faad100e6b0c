#!/usr/bin/env python3
"""Instruction mix of a kernel by what the instructions are FOR, from an .ncu-rep captured with --import-source on:

    python tools/ncu_categories.py gpurun_out/prof.ncu-rep initial_kernel [git-revision-of-the-capture]

Every CUDA-C source line of the kernel (ncu --page source --print-source sass,cuda) is assigned to a category by the file
it lives in and its line range (functions of device_common.cuh / reservoir.cuh / romis_detmath.h / romis_rng.h), and the
warp instructions executed are summed per category."""
import collections, csv, io, re, subprocess, sys, os

rep, kern = sys.argv[1], sys.argv[2]
REV = sys.argv[3] if len(sys.argv) > 3 else None          # git revision the capture was taken at (line numbers must match)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def source(path):
    if REV:
        return subprocess.run(["git", "-C", ROOT, "show", REV + ":" + os.path.relpath(path, ROOT)], capture_output=True, text=True).stdout
    return open(path).read()


def func_ranges(path, names):
    """line ranges of the named functions (from the line that holds the name followed by '(' to the closing brace at column 0/4)."""
    out = []
    lines = source(path).split("\n")
    for name, cat in names:
        for i, l in enumerate(lines):
            if re.search(r"\b" + re.escape(name) + r"\s*\(", l) and ("__device__" in l or "ROMIS_HD" in l or "ROMIS_RNG_HD" in l or "template" in lines[i - 1]):
                depth = 0; j = i; seen = False
                while j < len(lines):
                    depth += lines[j].count("{") - lines[j].count("}")
                    if "{" in lines[j]: seen = True
                    if seen and depth <= 0: break
                    j += 1
                out.append((i + 1, j + 1, cat)); break
    return out


DC = os.path.join(ROOT, "romis_b200", "csrc", "device_common.cuh")
RS = os.path.join(ROOT, "romis_b200", "csrc", "reservoir.cuh")
CATS = {
    "device_common.cuh": func_ranges(DC, [
        ("add3", "fp32 vector arithmetic (GLM order, no FMA)"), ("sub3", "fp32 vector arithmetic (GLM order, no FMA)"), ("mul3", "fp32 vector arithmetic (GLM order, no FMA)"),
        ("scale3", "fp32 vector arithmetic (GLM order, no FMA)"), ("dot3", "fp32 vector arithmetic (GLM order, no FMA)"), ("cross3", "BVH traversal (boxes, triangles, stack)"),
        ("mix3", "light sample (record fetch, position / colour)"), ("div3", "IEEE division / sqrt / reciprocal"), ("length3", "IEEE division / sqrt / reciprocal"),
        ("normalize3", "IEEE division / sqrt / reciprocal"), ("anynan3", "Phong evaluation (geometry, lobe test, NaN rules)"),
        ("light_sample", "light sample (record fetch, position / colour)"), ("tri_test", "BVH traversal (boxes, triangles, stack)"),
        ("box_test", "BVH traversal (boxes, triangles, stack)"), ("slab_inv", "BVH traversal (boxes, triangles, stack)"), ("load_node", "BVH traversal (boxes, triangles, stack)"),
        ("trace_closest", "BVH traversal (boxes, triangles, stack)"), ("trace_any", "BVH traversal (boxes, triangles, stack)"),
        ("gen_ray_dir", "pixel context (G-buffer, material, camera ray)"), ("diffuse_albedo", "pixel context (G-buffer, material, camera ray)"),
        ("make_ctx", "pixel context (G-buffer, material, camera ray)"), ("compute_shading", "Phong evaluation (geometry, lobe test, NaN rules)"),
        ("target_pdf", "Phong evaluation (geometry, lobe test, NaN rules)"), ("visible", "shadow-ray set-up"), ("tone_map", "tone mapping"),
        ("thread_pixel", "pixel context (G-buffer, material, camera ray)")]),
    "k_rmis.cu": func_ranges(os.path.join(ROOT, "romis_b200", "csrc", "k_rmis.cu"), [
        ("lemire32", "selection sampling (libstdc++ std::sample replay)"), ("emit_class", "selection sampling (libstdc++ std::sample replay)"),
        ("are_similar", "window classification (areSimilar)")]),
    "reservoir.cuh": func_ranges(RS, [
        ("res_init", "reservoir update / finish / store"), ("res_update", "reservoir update / finish / store"), ("res_store", "reservoir update / finish / store"),
        ("res_held_pdf", "reservoir update / finish / store"), ("res_finish", "reservoir update / finish / store"), ("stream_sample", "reservoir update / finish / store"),
        ("res_take_counts", "reservoir update / finish / store"), ("sat_add_u32", "reservoir update / finish / store")]),
}
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
h = rows[hi[0]]
iI, iT = h.index("Instructions Executed"), h.index("Thread Instructions Executed")
agg = collections.Counter(); thr = collections.Counter()
cur = ""
for r in rows[:hi[0]]:
    if r and r[0] in ("File Name", "File Path"): cur = r[1].split("/")[-1]
for r in rows[hi[0] + 1:]:
    if len(r) < len(h) or not r[0] or r[0] == "Line No":
        if r and r[0] in ("File Name", "File Path"): cur = r[1].split("/")[-1]
        continue
    try:
        ln, inst, t = int(r[0]), int(r[iI] or 0), int(r[iT] or 0)
    except ValueError:
        continue
    cat = None
    if cur == "romis_detmath.h": cat = "pow / exp in binary64 (deterministic)"
    elif cur == "romis_rng.h": cat = "counter-based random draws"
    elif cur in CATS:
        for a, b, c in CATS[cur]:
            if a <= ln <= b: cat = c; break
    if cat is None:
        cat = "kernel body (loop control, record loads / stores, light pick)" if cur.startswith("k_") else "other (" + cur + ")"
    agg[cat] += inst; thr[cat] += t
tot = sum(agg.values())
print(f"### `{kern}`: {tot / 1e6:.0f} M warp instructions (all captured launches), by purpose\n")
print("| purpose | share of warp instructions | active threads / warp inst |\n|---|---|---|")
for c, v in agg.most_common():
    if v and 100 * v / tot >= 0.05: print(f"| {c} | {100 * v / tot:.1f} % | {thr[c] / v:.1f} |")
