import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tools')
import numpy as np
from proto_shooting import *
from oracle import ballooning_oracle as bo
from ideal_ballooning_solver_b200 import synthetic

def solve(P, tol=2.0**-50, maxit=80, margin=0.0, T=32):
    lo, hi = P['Lb'], P['U']
    scale = max(abs(lo), abs(hi))
    lam = hi; nev=0; hist=[]
    above = []      # list of (lam, N) for points above lam1 (count 0), most recent last
    final=None
    for it in range(maxit):
        E = chunk_eval(P, lam, T=T); nev+=1
        rho = E['rho']; hist.append((lam,E['count']))
        inbasin = (E['nodes']==0 and E['r']>0)
        if E['count']==0:
            hi = min(hi, lam); above.append((lam, lam-rho))
            lo = max(lo, rho)
        else:
            lo = max(lo, lam)
            if inbasin: lo = max(lo, rho)
        if E['nodes']==0 and abs(rho-lam) <= tol*scale:
            final = rho; break
        if hi-lo <= tol*scale:
            final = 0.5*(lo+hi); break
        if inbasin:
            nxt = rho
        elif E['count']==0:
            if len(above)>=2:
                (b1,N1),(b2,N2) = above[-2], above[-1]
                den = N1-N2
                p = (b1-b2)/den if den>0 else 1.0
                p = min(1.0, max(0.4, p))
            else:
                p = 0.5
            b2,N2 = above[-1]
            nxt = b2 - p*N2*(1-margin)
            if not (lo < nxt < hi): nxt = 0.5*(lo+hi)
        else:
            nxt = 0.5*(lo+hi)
        if nxt==lam: final=lam; break
        lam = nxt
    return (final if final is not None else lam), nev, hist

def spectrum(theta,g,c,f):
    from scipy.linalg import eigh_tridiagonal
    hh, gu, cu, fu, sub, diag, sup = bo.discretise(theta, g, c, f)
    fi = fu[1:-1]; e = sup*np.sqrt(fi[:-1]/fi[1:])
    return eigh_tridiagonal(diag, e, eigvals_only=True, select='i', select_range=(len(diag)-3,len(diag)-1))

def problems():
    G = np.load('tests/golden/s_alpha.npz')
    theta = G['theta_1024']
    for ci,(sh,al,t0) in enumerate(G['cases']):
        g,c,f = synthetic.s_alpha_coefficients(sh,al,t0,theta)
        yield 'salpha%d'%ci, theta, g,c,f
    for name in ['ncsx_wout_op','synthetic_ncsx','synthetic_d3d','synthetic_hberg']:
        D = np.load('tests/golden/%s.npz'%name); theta=D['theta']
        ns,na,nl = D['geo_bmag'].shape
        for i in range(ns):
            for j in range(na):
                for k,th0 in enumerate(D['theta0s']):
                    cv = D['geo_cvdrift'][i,j]+th0*D['geo_cvdrift0'][i,j]
                    gd = D['geo_gds2'][i,j]+2*th0*D['geo_gds21'][i,j]+th0**2*D['geo_gds22'][i,j]
                    g,c,f = bo.gcf(D['dPdrho'][i,j], D['geo_bmag'][i,j], D['geo_gradpar_theta_pest'][i,j], cv, gd)
                    yield '%s_%d_%d_%d'%(name,i,j,k), theta, g,c,f

if __name__=='__main__':
    margin = float(sys.argv[1]) if len(sys.argv)>1 else 0.0
    tot=0; n=0; worst=0
    for name,theta,g,c,f in problems():
        h=theta[1]-theta[0]
        P = setup(g,c,f,h)
        la = spectrum(theta,g,c,f)
        lam, nev, hist = solve(P, margin=margin)
        tot+=nev; n+=1; worst=max(worst,nev)
        print('%-26s nev %2d err %9.2e gap %8.2e scale %7.1f'%(name,nev,lam-la[-1],la[-1]-la[-2],max(abs(P['Lb']),abs(P['U']))), [(float('%.4g'%a),b) for a,b in hist][:9])
    print('mean nev', tot/n, 'worst', worst)
