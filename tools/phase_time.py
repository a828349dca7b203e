"""Wall-clock per pipeline phase (with a sync after each), to find host-side overheads."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hiccup_b200 import _lib
from hiccup_b200.batch import DctBatchCodec
import bench
n,h,w=(int(a) for a in sys.argv[1:4]) if len(sys.argv)>=4 else (1024,426,640)
codec=DctBatchCodec(n,h,w)
base=bench.synthetic_batch(min(n,64),h,w,2000)
rgb=np.concatenate([base]*(n//len(base)+1))[:n]
codec.upload(rgb); _lib.sync()
lib=codec.lib
def T(name,fn,acc):
    t=time.perf_counter(); r=fn(); _lib.sync(); acc.setdefault(name,[]).append((time.perf_counter()-t)*1e3); return r
acc={}
for it in range(4):
    T('forward',lambda:_lib.check(lib.hic_dct_forward(codec.d_rgb.ptr,n,h,w,codec.d_coef.ptr,codec.d_ties.ptr,codec.blocks,codec.d_stats.ptr,None)),acc)
    T('symbolize',lambda:codec.encoder.symbolize(codec.d_coef.ptr,None),acc)
    T('build_codes',lambda:codec.encoder.build_codes(None),acc)
    T('pack',lambda:codec.encoder.pack(None),acc)
    tabs=T('tables()',lambda:codec.encoder.tables(),acc)
    enc=codec.encoder
    T('set_tables',lambda:_lib.check(lib.hic_decode_set_tables(codec.decoder.plan, enc.rows.ctypes.data, tabs[0].ctypes.data, tabs[1].ctypes.data, tabs[2].ctypes.data, None)),acc)
    T('decode_run',lambda:_lib.check(lib.hic_decode_run(codec.decoder.plan, enc._out.ptr, enc.byte_off.ctypes.data, enc.nbits.ctypes.data, codec.d_coef_dec.ptr, None)),acc)
    T('inverse',lambda:_lib.check(lib.hic_dct_inverse(codec.d_coef_dec.ptr,n,h,w,codec.d_y.ptr,codec.d_cr.ptr,codec.d_cb.ptr,codec.d_out.ptr,codec.d_ties.ptr,codec.blocks,codec.d_stats.ptr,None)),acc)
for k,v in acc.items(): print('%-12s %8.2f ms (min of %d)'%(k,min(v[1:]),len(v)-1))
print('rows total', enc.total_rows, 'bytes', enc.total_bytes)
