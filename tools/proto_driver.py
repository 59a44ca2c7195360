import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tools')
import numpy as np
from proto_shooting import *
from oracle import ballooning_oracle as bo
from ideal_ballooning_solver_b200 import synthetic

def trial_rq(P, theta):
    N=P['N']; span = theta[-1]
    v = np.zeros(N); th=theta[1:-1]
    v[1:-1] = (1-np.tanh(th/np.pi)**2)*np.cos(th/(2*span/np.pi))
    num = -np.sum(P['gh']*(v[1:]-v[:-1])**2) + np.sum(P['C']*v*v)
    return num/np.sum(P['F']*v*v)

def solve(P, theta, strategy='A', tol=4e-16, maxit=60, verbose=False):
    lo, hi = P['Lb'], P['U']
    nev = 0
    scale = max(abs(lo), abs(hi))
    lam = hi
    hist=[]
    best=None
    if strategy=='B':
        lo = max(lo, trial_rq(P, theta))
        lam = lo
    for it in range(maxit):
        E = chunk_eval(P, lam); nev += 1
        rho = E['rho']
        hist.append((lam, E['count'], rho))
        lo_prev = lo
        if E['count']==0: hi = min(hi, lam)
        else: lo = max(lo, lam)
        if np.isfinite(rho): lo = max(lo, min(rho, hi))
        # convergence: positive vector (nodes==0) and tiny correction
        if E['nodes']==0 and abs(rho-lam) <= tol*scale + 1e-15*abs(lam)*0:
            best=(rho, E); break
        if hi-lo <= tol*scale:
            best=(lam,E); break
        # next shift
        if E['nodes']==0 and E['r']>0 and rho>lam:
            nxt = rho            # in the basin (p12, lam1): Newton/RQI is monotone from below
        elif E['count']==0 and rho > lo_prev:
            nxt = rho            # from above: try the Rayleigh quotient (undershoots)
        else:
            nxt = 0.5*(lo+hi)
        if nxt==lam: nxt = 0.5*(lo+hi)
        lam = nxt
    return lam, nev, hist

if __name__=='__main__':
    from scipy.linalg import eigh_tridiagonal
    G = np.load('tests/golden/s_alpha.npz')
    theta = G['theta_1024']; h = theta[1]-theta[0]
    for strat in ['A','B']:
        tot=0
        for ci,(sh,al,t0) in enumerate(G['cases']):
            g,c,f = synthetic.s_alpha_coefficients(sh,al,t0,theta)
            P = setup(g,c,f,h)
            hh, gu, cu, fu, sub, diag, sup = bo.discretise(theta, g, c, f)
            fi = fu[1:-1]; e = sup*np.sqrt(fi[:-1]/fi[1:])
            lam_all = eigh_tridiagonal(diag, e, eigvals_only=True)
            lam, nev, hist = solve(P, theta, strat)
            tot+=nev
            print(strat, ci, 'nev',nev,'err',lam-lam_all[-1], 'gap', lam_all[-1]-lam_all[-2], [ (round(a,6),b) for a,b,_ in hist][:12])
        print('total', tot)
