"""Prove that a kernel source edit left existing kernels' device code unchanged: compile one .cu of
krypy_b200/csrc at a git revision and in the working tree (same nvcc flags as the Makefile) and
compare the SASS of every kernel that exists in both, instruction text only (addresses within a
function are relative; symbol names are normalised, an extra trailing template argument is allowed).

usage: python tools/sass_identity.py <git-rev> kry_orth.cu
       -> e.g. "orth_kernel<double,2,false>: SAME (9876 lines)"; exit code 1 on any difference

Used for the JT template parameter of orth_kernel: the instantiations that
were validated on the B200 are byte-identical after the edit, so the default path's device code is
the validated one."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "krypy_b200", "csrc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler",
         "-fPIC,-fvisibility=default"]


def sass(obj):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
        elif cur is not None:
            out[cur].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).rstrip())
    return out


def demangle(name):
    return subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name


def main():
    rev, src = sys.argv[1], sys.argv[2]
    tmp = tempfile.mkdtemp()
    old_dir = os.path.join(tmp, "krypy_b200", "csrc")
    os.makedirs(old_dir)
    os.makedirs(os.path.join(tmp, "include"))
    for rel in ["krypy_b200/csrc/" + f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))] + \
            ["include/krypy_b200.h"]:
        res = subprocess.run(["git", "-C", ROOT, "show", "%s:%s" % (rev, rel)], capture_output=True)
        if res.returncode == 0:
            with open(os.path.join(tmp, rel), "wb") as f:
                f.write(res.stdout)
    objs = []
    for d in (old_dir, CSRC):
        obj = os.path.join(tmp, "new.o" if d == CSRC else "old.o")
        subprocess.run(["nvcc"] + FLAGS + ["-c", os.path.join(d, src), "-o", obj], check=True, cwd=d,
                       stderr=subprocess.DEVNULL)
        objs.append(obj)
    old, new = sass(objs[0]), sass(objs[1])
    # default template arguments appended by later edits (old kernel == new kernel with these defaults)
    DEFAULTS = [", (int)16, (bool)0>", ", (int)16>"]
    dn_new = {n: demangle(n) for n in new}
    bad = 0
    matched = set()
    for name, body in sorted(old.items()):
        dn = demangle(name)
        cands = [n for n, d in dn_new.items() if d == dn or any(d == dn.replace(">(", suf + "(", 1) for suf in DEFAULTS)]
        if not cands:
            print("%s: not present any more" % dn)
            bad += 1
            continue
        n2 = cands[0]
        matched.add(n2)
        same = [ln.replace(name, "F") for ln in body] == [ln.replace(n2, "F") for ln in new[n2]]
        print("%s: %s (%d lines)" % (dn, "SAME" if same else "DIFFERENT", len(body)))
        bad += not same
    for n in sorted(set(new) - matched):
        print("new kernel: %s" % dn_new[n])
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
