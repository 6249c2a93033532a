"""Parity sweep run by tools/sanitize_host.sh under ASan + UBSan (test-only host simulator build)."""
import os
import sys
import tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import vs_testlib as T, numpy as np
T.HOSTSIM_SO='/tmp/libvsgpu_hostsim_asan.so'
from variantstore_b200 import VariantStoreIndex, load_library
lib=load_library(T.HOSTSIM_SO, subset=True)
for overlap,sparse in [(False,False),(True,False),(False,True),(True,True)]:
    for seed in range(2):
        with tempfile.TemporaryDirectory() as tmp:
            fa,vcf,names=T.write_fuzz_inputs(tmp,seed,overlap=overlap,sparse=sparse)
            o=T.Oracle.construct(fa,vcf,tmp+"/ser",force_enc=0 if sparse else -1)
            for cache in (None, tmp+"/cache"):
                if cache: os.makedirs(cache,exist_ok=True); os.environ["VSGPU_INDEX_CACHE"]=cache
                else: os.environ.pop("VSGPU_INDEX_CACHE",None)
                for rep in range(2 if cache else 1):
                    e=VariantStoreIndex(tmp+"/ser", lib=lib)
                    x,y,s=T.random_regions(seed+100,400,4000,n_samples=len(names))
                    b6,b4,_=T.compare_all(o,e,x,y,s)
                    x[:6]=np.arange(6); y[:3]=[0,2**40,1]
                    b2,_=T.compare_t2(o,e,x,y,s); b3,_=T.compare_t3(o,e,x,y,s); b5,_=T.compare_t5(o,e,x,y,s)
                    b1=T.compare_t1(o,e,x[6:])
                    assert not (b6 or b4 or b2 or b3 or b5 or b1), (overlap,sparse,seed,b6[:3],b4[:3],b2[:3],b3[:3],b5[:3],b1[:3])
                    rows=e.get_sample_var_in_sample(1,4001,names[0]); e.get_var_in_ref(1,4001)
                    e.close()
print("asan/ubsan run clean")
