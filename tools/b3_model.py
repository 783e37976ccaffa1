"""FP64 numpy model of the stage factorisation used by csrc/fused_b3.cu (prints the relative error against numpy.fft)."""
import numpy as np
N=128
rng=np.random.default_rng(0)
x=rng.normal(size=(N,N))+1j*rng.normal(size=(N,N))   # x[z][y]
F=rng.normal(size=(N,N))+1j*rng.normal(size=(N,N))   # F[kz][ky]
W=lambda n,e: np.exp(2j*np.pi*e/n)
# reference: P over both axes (exp(+)), multiply, P again
def P2(a):  # unnormalised DFT with exp(+2 pi i)
    return np.fft.ifft2(a)*a.size
ref=P2(P2(x)*F)   # ref[z][y]
# ---- model
ev=x[:,0::2]; od=x[:,1::2]           # [z][c], c<64
def s1(a):   # a[z][c] -> Y[z][c0][k']
    Y=np.zeros((N,2,32),complex)
    for c0 in range(2):
        sub=a[:,c0::2]               # [z][r]
        for k in range(32):
            Y[:,c0,k]=(sub*W(32,np.arange(32)*k)[None,:]).sum(1)
    return Y
def s2a(Y):
    k=np.arange(32)
    Y1=Y[:,1,:]*W(64,k)[None,:]
    X=np.zeros((N,64),complex); X[:,:32]=Y[:,0,:]+Y1; X[:,32:]=Y[:,0,:]-Y1
    return X
E=s2a(s1(ev)); O=s2a(s1(od))
assert np.allclose(E, np.fft.ifft(ev,axis=1)*64)
k=np.arange(64)
T=np.zeros((N,N),complex); T[:,:64]=E+W(128,k)[None,:]*O; T[:,64:]=E-W(128,k)[None,:]*O
assert np.allclose(T, np.fft.ifft(x,axis=1)*128)
# z forward: Z1[w][k1][col]
Z1=np.zeros((8,16,N),complex)
for w in range(8):
    rows=T[w::8,:]   # n1 index
    for k1 in range(16):
        Z1[w,k1,:]=(rows*W(16,np.arange(16)*k1)[:,None]).sum(0)*W(128,w*k1)
Zf=np.zeros((N,N),complex)
for k1 in range(16):
    for k0 in range(8):
        Zf[k1+16*k0,:]=(Z1[:,k1,:]*W(8,np.arange(8)*k0)[:,None]).sum(0)
assert np.allclose(Zf, P2(x))
Pm=Zf*F
G=np.zeros((8,16,N),complex)
for zl in range(8):
    for k1 in range(16):
        G[zl,k1,:]=sum(Pm[k1+16*k0,:]*W(8,k0*zl) for k0 in range(8))*W(128,k1*zl)
H=np.zeros((N,N),complex)
for zl in range(8):
    for n1 in range(16):
        H[zl+8*n1,:]=sum(G[zl,k1,:]*W(16,k1*n1) for k1 in range(16))
u=H[:,:64]+H[:,64:]; v=(H[:,:64]-H[:,64:])*W(128,k)[None,:]
def inv64(a):  # a[z][k] -> out[z][m]
    out=np.zeros((N,64),complex)
    kp=np.arange(32)
    for m0 in range(2):
        p=(a[:,:32]+(-1)**m0*a[:,32:])*W(64,kp*m0)[None,:]
        for m1 in range(32):
            out[:,m0+2*m1]=(p*W(32,kp*m1)[None,:]).sum(1)
    return out
y=np.zeros((N,N),complex); y[:,0::2]=inv64(u); y[:,1::2]=inv64(v)
print("max err", np.abs(y-ref).max()/np.abs(ref).max())
