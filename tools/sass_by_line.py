"""Static SASS instruction counts of one kernel by source function (needs -lineinfo).
usage: sass_by_line.py <cubin> <kernel-substring> <source.h> [more sources...]
Functions are recognised by 'QPB_HD|__device__|template' definitions in the given sources (line ranges)."""
import collections, re, subprocess, sys

def func_ranges(path):
    out = []
    lines = open(path).read().split('\n')
    name = None
    for i, l in enumerate(lines, 1):
        m = re.match(r'^(?:QPB_HD|__device__ __forceinline__|inline|__global__)[^(]*?\b(\w+)\(', l) or re.match(r'^(\w+)\(const __grid_constant__', l)
        if m:
            name = m.group(1); out.append([name, i, None])
        if l.startswith('}') and out and out[-1][2] is None:
            out[-1][2] = i
    return out

def main():
    cubin, kern = sys.argv[1], sys.argv[2]
    srcs = sys.argv[3:]
    rngs = {s.split('/')[-1]: func_ranges(s) for s in srcs}
    txt = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
    infun = False; cur = None
    cnt = collections.Counter(); ops = collections.defaultdict(collections.Counter)
    for line in txt.split('\n'):
        if line.startswith('//---') and '.text.' in line:
            infun = kern in line
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
        if not infun or cur is None: continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)', line)
        if m:
            f, l = cur
            fn = f
            for name, a, b in rngs.get(f, []):
                if b and a <= l <= b: fn = f + ':' + name
            cnt[fn] += 1; ops[fn][m.group(1)] += 1
    tot = sum(cnt.values())
    print('total', tot)
    for fn, c in cnt.most_common():
        top = ', '.join(f'{o} {k}' for o, k in ops[fn].most_common(6))
        print(f'{c:6d}  {fn:40s} {top}')

main()
