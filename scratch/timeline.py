import sys, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests/golden')
import bench
m = bench.make_model(1000000, 0)
e = m._engine
e.lib.mrgp_timeline_enable(e.handle, 1)
for s in range(12): e.sweep(1)
e.synchronize()
tags=(C.c_int32*128)(); ms=(C.c_float*256)()
n=e.lib.mrgp_timeline_read(e.handle, tags, ms, 128)
J=10; n=4*J
names=['A','mid','B','omega']
for j in range(n//4):
    line='L%d:'%j
    for k in range(4):
        b,en=ms[2*(4*j+k)]*1000, ms[2*(4*j+k)+1]*1000
        line+='  %s [%.1f -> %.1f] %.1f'%(names[k], b, en, en-b)
    print(line)

for j in range(J):
    line='L%d omega detail:'%j
    for k,nm in enumerate(['ard','prologue','iterate']):
        q=48+4*j+k
        b,en=ms[2*q]*1000, ms[2*q+1]*1000
        line+='  %s [%.1f -> %.1f] %.1f'%(nm,b,en,en-b)
    print(line)
