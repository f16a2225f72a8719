import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, _oracle as O, _problems as P, pgopt_b200 as pg
basis, om, A = pg.KnownBasisVB(), O.known_model(), P.plausible_A_known()
for (N,T) in [(30,2),(30,3),(30,10),(1500,3)]:
    u,y,_=P.make_data(T,11); s=P.injected_streams(N,T,2,5,21)
    xp=np.random.default_rng(4).normal(2,1,(2,T))
    st=O.make_streams(philox=False, **{k:s[k] for k in ("Z0","Ures","Z","Uanc","Ustar")})
    ref=O.sweep(om,N,T,u,y,[[1.0,0.0]],0.1,[2.0,2.0],np.eye(2),A,P.Q_TRUE,xp,False,st,2)
    e=pg.ParticleGibbsEngine(basis,1,N,T,keep_w_history=True)
    e.set_data(u,y); e.set_measurement(pg.LinearMeasurement([[1.0,0.0]]),0.1); e.set_init_dist(pg.MvNormal([2.0,2.0],1.0))
    e.set_state(A,P.Q_TRUE,xp,2); print("after set_state status",e.status())
    e.set_rng_injected(s["Z0"],s["Ures"],s["Z"],s["Uanc"],s["Ustar"])
    e.sweep_filter()
    a,x,w=e.get_history()
    print(N,T,"status",e.status(),"ref",ref["status"],"a eq",np.array_equal(a[1:],ref["a"][1:]),"x eq",np.array_equal(x,ref["x"]),"w eq",np.array_equal(w,ref["w"]),"nan x",np.isnan(x).sum(),"nan w",np.isnan(w).sum())
    if not np.array_equal(a[1:],ref["a"][1:]):
        bad=np.argwhere(a[1:]!=ref["a"][1:]); print(" first bad a", bad[:5], a[1:][tuple(bad[0])], ref["a"][1:][tuple(bad[0])])
    e.close()
