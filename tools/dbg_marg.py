import numpy as np, sys
sys.path.insert(0, '/root/repo')
from dsopp_b200 import capi, synth
from oracle import pba_oracle as O
SIGMA=20.0
win = synth.make_window(n_frames=4, points_per_frame=200, seed=32, ab_scale=0.0)
rng = np.random.default_rng(0)
A = rng.normal(size=(64, 32)); Hm = A.T @ A * 50.0; bm = rng.normal(size=32) * 5.0
frames = O.frames_from_window(win); O.first_estimate_jacobians(frames)
trace=[]
e_ref, _, _ = O.lm_solve(O.Problem(frames, SIGMA, Hm, bm, 12.5), O.LMOptions(7, 1e-5, 1e-8, 1e-8, True, 3, 1.0, 1.0), trace)
for g in (1,0):
    h = capi.upload_window(win); h.set_option("cuda_graph", g) if hasattr(h,'set_option') else None
    h.first_estimate()
    e, it, _, _ = h.solve_lm(SIGMA, H_marg=Hm, b_marg=bm, energy_marg=12.5)
    s, _ = h.get_state()
    ref = O.state_eps_stacked(frames)
    print("graph",g,"E",e,e_ref,"it",it,len(trace))
    print(np.abs(s-ref).reshape(4,8))
    print(ref.reshape(4,8))
    h.close()
