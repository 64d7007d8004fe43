import sys, time, os
t0=time.perf_counter()
import cv2, numpy as np
t1=time.perf_counter()
if len(sys.argv)>1 and sys.argv[1]=="noocl": cv2.ocl.setUseOpenCL(False)
if len(sys.argv)>2: cv2.setNumThreads(int(sys.argv[2]))
rng=np.random.default_rng(0)
img=rng.integers(0,256,(240,320,3),dtype=np.uint8)
t2=time.perf_counter(); f=cv2.bilateralFilter(img,9,75,75); t3=time.perf_counter()
s=cv2.SIFT_create(); t4=time.perf_counter(); kp,d=s.detectAndCompute(f,None); t5=time.perf_counter()
kp,d=s.detectAndCompute(f,None); t6=time.perf_counter()
from concurrent.futures import ThreadPoolExecutor
def one(i): return cv2.SIFT_create().detectAndCompute(cv2.bilateralFilter(img,9,75,75),None)[1].shape
t7=time.perf_counter()
with ThreadPoolExecutor(16) as p: r=list(p.map(one, range(32)))
t8=time.perf_counter()
with ThreadPoolExecutor(16) as p: r=list(p.map(one, range(32)))
t9=time.perf_counter()
print(sys.argv[1:], "import %.2f bilateral1 %.3f create %.3f sift1 %.3f sift2 %.3f pool1 %.3f pool2 %.3f"%(t1-t0,t3-t2,t4-t3,t5-t4,t6-t5,t8-t7,t9-t8))
