"""SASS instruction count per source function of the align kernel: nvdisasm -g -c <cubin> output on stdin-file,
line markers mapped to the enclosing function of c2b_core.cuh / c2b_engine.cu by line ranges.
usage: code_by_function.py dis.txt crispresso2_b200/csrc/c2b_core.cuh"""
import collections
import re
import sys

dis, core = sys.argv[1], sys.argv[2]
funcs = []                                  # (start_line, name)
for k, ln in enumerate(open(core), 1):
    m = re.match(r"^(?:template.*\n)?(?:C2B_DEVNOINL|C2B_DEV)\s+[\w:<> \*&]+?\s+(\w+)\(", ln)
    if m:
        funcs.append((k, m.group(1)))
starts = [f[0] for f in funcs]


def func_of(line):
    import bisect
    i = bisect.bisect_right(starts, line) - 1
    return funcs[i][1] if i >= 0 else "?"


cnt = collections.Counter()
cur = ("?", 0)
inl = None
for ln in open(dis):
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        f = m.group(1).split("/")[-1]
        cur = (f, int(m.group(2)))
        continue
    if re.match(r"^\s+/\*[0-9a-f]{4,6}\*/", ln):
        name = func_of(cur[1]) if cur[0] == "c2b_core.cuh" else cur[0]
        cnt[name] += 1
tot = sum(cnt.values())
for k, v in cnt.most_common(40):
    print("%6d %5.1f%%  %s" % (v, 100.0 * v / tot, k))
print(tot)
