"""numpy prototype of the far-field device math (affine-map formulation); dev tool only."""
import numpy as np
A1=np.zeros((7,7)); A1[0]=[0,0,-57.74022,1.5373825,1.154294,-0.441762,0.0906601]
A1[2]=[-72.17349,1.8691899,2.8859421,-1.026471,0.2610568,0,0]; A1[4]=[1.8083973,-0.603195,0.2177414,0,0,0,0]; A1[6,0]=0.0394559
A2=np.zeros((8,8)); A2[0]=[0,0,103.90461,6.6513025,2.8405781,-0.7819705,-0.0400483,0.0896645]
A2[2]=[115.44758,7.3024355,5.7640389,-1.578144,-0.0354326,0.2781226,0,0]
A2[4]=[2.9130983,-0.8104051,-0.0185283,0.2626023,0,0,0,0]; A2[6,:2]=[-0.0250794,0.0709672]
RN=3000.
def Rx(a): c,s=np.cos(a),np.sin(a); return np.array([[1,0,0],[0,c,-s],[0,s,c]])
def Ry(a): c,s=np.cos(a),np.sin(a); return np.array([[c,0,s],[0,1,0],[-s,0,c]])
def Rea(el,az): return Ry(az)@Rx(el)
def poly(A,x,y):
    u=x/RN; v=y/RN; n=A.shape[0]
    z=0; zx=0; zy=0
    for i in range(n):
        for j in range(n):
            if A[i,j]==0: continue
            z=z+A[i,j]*u**i*v**j
            if i>0: zx=zx+A[i,j]*(i/RN)*u**(i-1)*v**j
            if j>0: zy=zy+A[i,j]*u**i*(j/RN)*v**(j-1)
    return z,zx,zy
def newton(A,p,d,niter,log):
    t=-p[2]/d[2]
    for it in range(niter):
        q=p+d*t
        z,zx,zy=poly(A,q[0],q[1])
        f=q[2]-z; fp=d[2]-(zx*d[0]+zy*d[1])
        dt=f/fp; t=t-dt; log.append(np.abs(dt).max())
    q=p+d*t; z,zx,zy=poly(A,q[0],q[1])
    N=np.array([-zx,-zy,np.ones_like(zx)]); N/=np.sqrt((N*N).sum(0))
    return t,q,N
def ff_trace(theta,phi,el,az,off_mm=1320.,src=(0,-7200.,1e8),niter=3):
    th,ph=np.meshgrid(theta,phi); th=th.ravel(); ph=ph.ravel()
    d=np.array([np.sin(th)*np.cos(ph),np.sin(th)*np.sin(ph),np.cos(th)])
    C=np.array([0,-7200.,0]); src=np.array(src,float)
    th1=np.arctan(1/2); th2=-np.pi/2-np.arctan(1/3)
    Rm=Rea(-el,-az); Rp=Rea(el,az)
    F1=np.diag([-1,1,-1.])@Rx(-th1); o1=np.array([0,7200.,3600.])
    F2=np.diag([-1,-1,1.])@Rx(-th2); o2=np.array([0,12000.,0])
    # aperture: (Rm(p-C)+C).z = 4000
    a=Rm[2]; s_ap=4000-C[2]-a@(src-C)
    t_ap=s_ap/(a@d)
    Pa=src[:,None]+d*t_ap
    valid=~(Pa[0]**2+(Pa[1]+7201.003729431267)**2>3e3**2)
    d=d[:,valid]; Pa=Pa[:,valid]; t_ap=t_ap[valid]
    T1A=F1@Rm; T1b=F1@(C-Rm@C+o1)
    T2A=F2@Rm; T2b=F2@(C-Rm@C+o2)
    B1=Rp@F1.T; B2=Rp@F2.T
    log=[]
    p1=T1A@Pa+T1b[:,None]; d1=T1A@d
    t1,h1,N1=newton(A1,p1,d1,niter,log)
    Pm1=Pa+d*t1
    d1r=d1-2*(d1*N1).sum(0)*N1
    dt2=B1@d1r
    p2=T2A@Pm1+T2b[:,None]; d2=T2A@dt2
    t2,h2,N2=newton(A2,p2,d2,niter,log)
    Pm2=Pm1+dt2*t2
    d2r=d2-2*(d2*N2).sum(0)*N2
    dt3=B2@d2r
    # focal: (Rp(p-C)+C).y = -off
    fa=Rp[1]; fb=C[1]-fa@C
    tf=(-off_mm-(fa@Pm2+fb))/(fa@dt3)
    pos=Pm2+dt3*tf
    opl=t_ap+t1+t2+tf
    # focal centre
    ax=np.linspace(-4000,4000,100); m=np.sqrt(ax**2+ax[50]**2)<3600
    row=np.array([ax[m],np.full(m.sum(),ax[50]*np.cos(np.pi/2)),np.full(m.sum(),ax[50])])
    col=np.array([np.full(m.sum(),ax[50]),ax[m]*np.cos(np.pi/2),ax[m]])
    rr=Rp@(row-C[:,None])+C[:,None]; cc=Rp@(col-C[:,None])+C[:,None]
    cen=np.array([rr[0].mean(),cc[1].mean(),cc[2].mean()])
    return valid,Pm1,Pm2,B2@N2,pos,dt3,opl,cen,log
if __name__=='__main__':
    G='/root/repo/tests/golden/'
    d=np.load(G+'ff_rays.npz')
    for i in range(int(d['n_cases'])):
        n=int(d['in_n_scan'][i]); th=np.linspace(np.pi-5e-5,np.pi+5e-5,n); ph=np.linspace(0,np.pi,n)
        rf=d['field_%d'%i]; rm=d['mirror_%d'%i]
        for niter in (2,3,4):
            valid,Pm1,Pm2,N2t,pos,dt3,opl,cen,log=ff_trace(th,ph,float(d['in_el'][i]),float(d['in_az'][i]),off_mm=float(d['in_offset'][i])*1e3,niter=niter)
            rv=rf[15]!=0
            print(i,niter,'mask',np.array_equal(valid,rv),'m2',np.abs(Pm2-rf[0:3][:,rv]).max(),'m1',np.abs(Pm1-rf[3:6][:,rv]).max(),'N',np.abs(N2t-rf[6:9][:,rv]).max(),
              'foc',np.abs(pos-cen[:,None]-rf[9:12][:,rv]).max(),'dir',np.abs(dt3-rf[12:15][:,rv]).max(),'opl ulp',np.abs(opl-rf[15][rv]).max()/1.4901161193847656e-08,'cen',np.abs(cen-rm[6:9,0]).max(),['%.1e'%x for x in log])
