"""Static SASS size of the bulk path of every k_step instantiation (no GPU needed):
    python scripts/sass_count.py [filter]
The kernel body is straight-line code, so the static count of the bulk path tracks the executed count.
The wall blocks (edge_block) precede the bulk path in the function; the bulk starts at the target of
the first uniform branch."""
import collections, os, re, subprocess, sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "vivsim_b200", "libvivsim_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
flt = sys.argv[1] if len(sys.argv) > 1 else ""
fn = None
funcs = collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1); funcs[fn] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and fn: funcs[fn].append((int(m.group(1), 16), m.group(2).strip()))
for fn, ins in funcs.items():
    if "k_step" not in fn: continue
    dem = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    m = re.search(r"k_step<\(int\)(\d), \(int\)(\d+), \(int\)(\d)>", dem)
    tag = "k_step<%s,%s,%s>" % m.groups() if m else dem[:40]
    if flt and flt not in tag: continue
    start = 0
    for a, s in ins:
        m = re.search(r"BRA\.U\s+!?UP\d, (0x[0-9a-f]+)", s)
        if m: start = int(m.group(1), 16); break
    body = [s for a, s in ins if a >= start]
    ops = collections.Counter()
    for s in body:
        t = s.split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops[op.split(".")[0]] += 1
    fl = ops["FADD"] + ops["FMUL"] + ops["FFMA"]
    print(f"{tag:16s} bulk {len(body):5d}  float {fl:5d}  LDL {ops['LDL']:4d} STL {ops['STL']:4d}  MUFU {ops['MUFU']:3d} "
          f"int {ops['IADD3']+ops['IMAD']+ops['LEA']+ops['ISETP']+ops['SEL']+ops['MOV']+ops['LOP3']+ops['SHF']:5d}  BRA {ops['BRA']:3d}")
