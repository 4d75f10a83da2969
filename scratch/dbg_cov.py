import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
src=open('/root/repo/tests/test_gpu_properties.py').read()
i=src.index("def _colliding_weight_map"); j=src.index("def test_coverage_pick_is_decided")
ns={'np':np,'torch':torch}; exec(src[i:j],ns)
vals,s,pairs=ns['_colliding_weight_map']()
from oracle import densify_oracle as O
from lichtfeld_densification_plugin_b200.core.sampling import select_samples_with_coverage
cert=torch.from_numpy(vals)
np.random.seed(5)
got=select_samples_with_coverage(cert,3000)
want=O.select_samples(cert,3000,rng=np.random.RandomState(5),s_override=s)
print(len(got),len(want))
a=set(got.tolist()); b=set(want.tolist())
print("only gpu",sorted(a-b)[:20]); print("only ref",sorted(b-a)[:20])
pi={i for i,_ in pairs}; pj={j for _,j in pairs}
print("gpu-only in pair-i:",len((a-b)&pi)," ref-only in pair-j:",len((b-a)&pj))
