"""How long does a process pay for the library?  b2d_init (CUDA context + streams), first call, shutdown, exit."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t00 = time.perf_counter()
import b2d_loader
b = b2d_loader.load()
t = time.perf_counter(); b.init(0); print(os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"), "init", round(time.perf_counter() - t, 3))
d = b.corpus("text", 1, 1 << 20)
t = time.perf_counter(); c = b.deflate_chunks(d); print("first deflate", round(time.perf_counter() - t, 3))
t = time.perf_counter(); b.shutdown(); print("shutdown", round(time.perf_counter() - t, 3))
print("total before exit", round(time.perf_counter() - t00, 3), flush=True)
