import sys, time, ctypes, zlib
sys.path.insert(0,'.')
import numpy as np, torch
import b2d_loader
b2d=b2d_loader.load(); b2d.init(0); L=b2d.lib()
n=1024; MB=256*1024
raw=np.empty(n*MB,np.uint8)
members=[]
for i in range(n):
    L.b2d_corpus_text(1000+i, raw[i*MB:].ctypes.data, MB)
    c=zlib.compressobj(1,zlib.DEFLATED,-15); members.append(c.compress(raw[i*MB:(i+1)*MB].data)+c.flush())
in_off=np.zeros(n+1,np.uint64); in_off[1:]=np.cumsum([len(m) for m in members]); tot=int(in_off[-1])
out_off=(np.arange(n+1,dtype=np.uint64)*MB)
h_blob=torch.empty(tot+64,dtype=torch.uint8).pin_memory(); h_blob[:tot]=torch.from_numpy(np.frombuffer(b"".join(members),np.uint8).copy())
h_out=torch.empty(n*MB,dtype=torch.uint8).pin_memory()
ol=np.zeros(n,np.uint64); ic=np.zeros(n,np.uint64); crc=np.zeros(n,np.uint32); st=np.zeros(n,np.int32)
def call(flags):
    t=time.perf_counter()
    r=L.b2d_inflate_batch(h_blob.data_ptr(), in_off.ctypes.data, n, h_out.data_ptr(), out_off.ctypes.data, ol.ctypes.data, ic.ctypes.data, crc.ctypes.data, st.ctypes.data, flags)
    assert r==0
    return (time.perf_counter()-t)*1e3
for f in (0,1,0,1,0,1): print('flags',f,'ms',round(call(f),2))
d=torch.empty(n*MB,dtype=torch.uint8,device='cuda')
torch.cuda.synchronize(); t=time.perf_counter(); d_blob=h_blob.to('cuda',non_blocking=True); torch.cuda.synchronize(); print('h2d ms',(time.perf_counter()-t)*1e3, tot/1e6,'MB')
t=time.perf_counter(); h_out.copy_(d,non_blocking=True); torch.cuda.synchronize(); print('d2h ms',(time.perf_counter()-t)*1e3)
t=time.perf_counter(); h_out.copy_(d,non_blocking=True); torch.cuda.synchronize(); print('d2h ms',(time.perf_counter()-t)*1e3)
# pinned via library
p=L.b2d_alloc_pinned(n*MB); 
t=time.perf_counter()
r=L.b2d_inflate_batch(h_blob.data_ptr(), in_off.ctypes.data, n, p, out_off.ctypes.data, ol.ctypes.data, ic.ctypes.data, crc.ctypes.data, st.ctypes.data, 0); print('lib-pinned out ms',(time.perf_counter()-t)*1e3)
t=time.perf_counter()
r=L.b2d_inflate_batch(h_blob.data_ptr(), in_off.ctypes.data, n, p, out_off.ctypes.data, ol.ctypes.data, ic.ctypes.data, crc.ctypes.data, st.ctypes.data, 0); print('lib-pinned out ms',(time.perf_counter()-t)*1e3)
