"""Per-kernel census of the Blackwell-specific SASS opcodes in the shipped library (no GPU needed):
    python tools/sass_census.py radar-camera-fusion-depth_b200/librcfd_b200.so > profiles/r2_sass_opcodes.txt"""
import re, subprocess, collections, sys
so = sys.argv[1]
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
ops = ['UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTCBAR', 'SYNCS', 'UCGABAR', 'MAPA']
cur, cnt = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    if cur:
        for o in ops:
            if re.search(r'\b' + o, line):
                cnt[cur][o] += 1
names = subprocess.run(['c++filt'], input='\n'.join(cnt.keys()), capture_output=True, text=True).stdout.splitlines()
rows = []
for (k, c), n in zip(cnt.items(), names):
    if not any(c[o] for o in ops):
        continue
    n = n.replace('void ', '').replace('rcfd::(anonymous namespace)::', '').replace('(anonymous namespace)::', '')
    n = re.sub(r'\(.*', '', n)
    rows.append((n, [c[o] for o in ops]))
rows.sort()
print('SASS opcode census of librcfd_b200.so (cuobjdump -sass, sm_100a cubin), per kernel: tcgen05.mma = UTCHMMA, tcgen05.ld / st = LDTM / STTM,')
print('TMA tile load = UTMALDG, tcgen05.commit / mbarrier = UTCBAR / SYNCS, cluster barrier = UCGABAR, distributed shared memory address = MAPA.')
print()
print('%-58s' % 'kernel' + ''.join('%9s' % o for o in ops))
for n, v in rows:
    print('%-58s' % n[:58] + ''.join('%9d' % x for x in v))
print()
print('%-58s' % 'total' + ''.join('%9d' % sum(v[i] for _, v in rows) for i in range(len(ops))))
