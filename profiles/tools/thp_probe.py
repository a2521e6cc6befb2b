import mmap, time, numpy as np, re
print(open('/sys/kernel/mm/transparent_hugepage/enabled').read().strip(), '|', open('/sys/kernel/mm/transparent_hugepage/defrag').read().strip())
def anon_huge():
    return int(re.search(r'AnonHugePages:\s+(\d+)', open('/proc/self/smaps_rollup').read()).group(1)) // 1024
n = 4 << 30
for adv in (False, True):
    m = mmap.mmap(-1, n)
    if adv: m.madvise(mmap.MADV_HUGEPAGE)
    a = np.frombuffer(m, np.uint8)
    t0 = time.perf_counter(); a[::4096] = 1; dt = time.perf_counter() - t0
    print('madvise', adv, f'touch {n/dt/1e9:.1f} GB/s, AnonHugePages {anon_huge()} MB')
    del a; m.close()
