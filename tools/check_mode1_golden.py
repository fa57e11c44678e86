import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from ofblend_b200 import capi, synth
from conftest import rel_l2
g=np.load('tests/golden/mode1_32x48.npz')
api=capi.HostAPI()
dims=tuple(int(x) for x in g['dims'])
i0 = synth.post_process(synth.two_drop_phi(dims, 0), api)
i1 = synth.post_process(synth.two_drop_phi(dims, 1), api)
v0 = np.zeros(i0.shape + (4,), np.float32)
vel, iters, errs = api.optical_flow_multiscale4d(v0, i0, i1, want_trace=True, **synth.MODE1_PARAMS)
print(iters, list(g['cg_iters']))
print(errs); print(list(g['errs']))
s=int(g['stride'])
print('rel', rel_l2(vel[::s,::s,::s,::s], g['vel_sub']), 'l2', np.linalg.norm(vel.astype(np.float64).ravel()), float(g['vel_l2']))
adv=api.advect4d(vel,i0)
print('sdf', np.abs(adv[::s,::s,::s,::s]-g['adv_sub']).max()/0.005)
