"""Source-level attribution of an ncu capture: python profiles/attribute.py <lib.so> <report.ncu-rep> <kernel-substring> [top]

ncu's CLI exports the per-instruction counters of a kernel only for the SASS view.  This script lines that listing up with
`nvdisasm -g` of the same cubin (same instruction order), which carries file:line for every instruction, and folds samples,
stall reasons, executed instructions and active lanes by (a) enclosing function of the source line and (b) source line.
The library must be the build that was profiled (-lineinfo)."""
import collections, csv, io, os, re, subprocess, sys, tempfile

lib, rep, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
kname = rows[0][1]
H = rows[1]
idx = {h: i for i, h in enumerate(H)}
data = [r for r in rows[2:] if len(r) >= len(H)]
# find the cubin function whose instruction count matches
mangled = None
lines_of = None
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"):
        continue
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur, cur_line, acc = None, None, {}
    for ln in dis.splitlines():
        m = re.match(r"^\.text\.(\S+):", ln)
        if m:
            cur = m.group(1); acc[cur] = []; cur_line = None
            continue
        m = re.match(r'^\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if cur and re.match(r"^\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            acc[cur].append(cur_line)
    for name, lst in acc.items():
        if kern in name and len(lst) == len(data):
            mangled, lines_of = name, lst
if lines_of is None:
    sys.exit("no cubin function matching %r with %d instructions" % (kern, len(data)))
# enclosing function per (file, line)
src_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "marl-mass_b200", "csrc")
func_at = {}
for fn in set(l[0] for l in lines_of if l):
    path = os.path.join(src_dir, fn)
    if not os.path.exists(path):
        continue
    cur = "?"
    for n, text in enumerate(open(path), 1):
        m = re.match(r"^(?:template\s*<[^>]*>\s*)?(?:static\s+)?__(?:device|global)__.*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text)
        if m and not text.startswith(" "):
            cur = m.group(1)
        func_at[(fn, n)] = cur
stalls = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
byf = collections.defaultdict(lambda: collections.Counter())
byl = collections.defaultdict(lambda: collections.Counter())
S = I = T = 0
for r, loc in zip(data, lines_of):
    s, i, t = int(r[idx["# Samples"]]), int(r[idx["Instructions Executed"]]), int(r[idx["Thread Instructions Executed"]])
    S += s; I += i; T += t
    fn = func_at.get(loc, "?") if loc else "?"
    for key, tab in ((fn, byf), (loc, byl)):
        c = tab[key]
        c["smp"] += s; c["inst"] += i; c["thr"] += t; c["code"] += 1
        for st in stalls:
            c[st] += int(r[idx[st]])
print("kernel %s\nsamples %d warp-inst %d lanes %.1f" % (kname, S, I, T / max(I, 1)))
tot = collections.Counter()
for c in byf.values():
    for st in stalls:
        tot[st] += c[st]
print(" ".join("%s %.1f%%" % (k[6:], 100.0 * v / S) for k, v in tot.most_common(8)))
for name, c in sorted(byf.items(), key=lambda kv: -kv[1]["smp"])[:top]:
    ts = sorted(((c[st], st[6:]) for st in stalls), reverse=True)[:3]
    print("%-26s smp %5.1f%% inst %5.1f%% lanes %4.1f code %5d  %s" % (name, 100.0 * c["smp"] / S, 100.0 * c["inst"] / I, c["thr"] / max(c["inst"], 1), c["code"],
                                                         " ".join("%s=%d%%" % (n, 100 * v / max(c["smp"], 1)) for v, n in ts)))
print("--- hottest source lines")
for loc, c in sorted(byl.items(), key=lambda kv: -kv[1]["smp"])[:top]:
    print("%-22s smp %5.1f%% inst %5.1f%% lanes %4.1f" % ("%s:%d" % loc if loc else "?", 100.0 * c["smp"] / S, 100.0 * c["inst"] / I, c["thr"] / max(c["inst"], 1)))
