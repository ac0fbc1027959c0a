import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from scipy import sparse
from deepsphere_tf1_b200 import utils, healpix as hp, graph_build
from deepsphere_tf1_b200._lib import call
from deepsphere_tf1_b200.graph import DeviceGraph
dev=torch.device('cuda',0)
nside=8
M=12*nside*nside
W=utils.healpix_weightmatrix(nside=nside,dtype=np.float32)
W.sort_indices()
coords=torch.empty((M,3),dtype=torch.float32,device=dev); call('hpx_coords', nside, None, M, coords)
nb=torch.empty((M,8),dtype=torch.int32,device=dev); d2=torch.empty((M,8),dtype=torch.float32,device=dev)
call('hpx_neighbours', nside, None, None, M, coords, nb, d2)
d2h=d2.cpu().numpy(); valid=nb.cpu().numpy()>=0
kw=np.mean(d2h[valid]); wh=np.exp(-d2h/(2*kw))
# reference weights in COO order
x,y,z=hp.pix2vec(nside, np.arange(M)); ch=np.vstack([x,y,z]).T.astype(np.float32)
nbh=hp.get_all_neighbours(nside, np.arange(M)).T
rows=np.repeat(np.arange(M),8).reshape(M,8)
dist=np.sum((ch[rows[valid]]-ch[nbh[valid]])**2,axis=1)
kw_ref=np.mean(dist); w_ref=np.exp(-dist/(2*kw_ref))
print('kw equal', kw==kw_ref, kw, kw_ref, type(kw))
print('w equal', np.array_equal(wh[valid], w_ref), np.abs(wh[valid]-w_ref).max())
w=torch.from_numpy(wh.astype(np.float32)).to(dev)
scol=torch.empty((M,8),dtype=torch.int32,device=dev); sw=torch.empty((M,8),dtype=torch.float32,device=dev)
cnt=torch.empty(M,dtype=torch.int32,device=dev); deg=torch.empty(M,dtype=torch.float32,device=dev)
call('hpx_sort_degree', nb, w, M, scol, sw, cnt, deg)
d_ref=np.ravel(W.sum(1))
print('deg equal', np.array_equal(deg.cpu().numpy(), d_ref), np.abs(deg.cpu().numpy()-d_ref).max(), d_ref.dtype)
# sorted weights equal CSR data?
swh=sw.cpu().numpy(); ch_=cnt.cpu().numpy()
flat=np.concatenate([swh[i,:ch_[i]] for i in range(M)])
print('sorted w equal csr data', np.array_equal(flat, W.data), 'cols', np.array_equal(np.concatenate([scol.cpu().numpy()[i,:ch_[i]] for i in range(M)]), W.indices))
d12=np.power(deg.cpu().numpy(),-0.5).astype(np.float32)
Lu=utils.build_laplacian(W,'normalized'); Lu=sparse.csr_matrix(Lu).copy(); Lu.sort_indices()
g2,_,_=graph_build.healpix_level(nside,None,dev,rescale=False)
ref2=DeviceGraph(Lu.copy(),dev)
a,b=g2.fwd[2].cpu().numpy(), ref2.fwd[2].cpu().numpy()
print('unscaled vals equal', np.array_equal(a,b), np.abs(a-b).max(), np.argwhere(a!=b)[:4].tolist())
lmax=np.float32(1.93)
Lh=utils.rescale_L(Lu.copy(),lmax=lmax).tocoo()
g,_,_=graph_build.healpix_level(nside,None,dev,lmax=lmax)
ref=DeviceGraph(Lh,dev)
a,b=g.fwd[2].cpu().numpy(), ref.fwd[2].cpu().numpy()
print('scaled vals equal', np.array_equal(a,b), np.abs(a-b).max(), np.argwhere(a!=b)[:4].tolist(), a[a!=b][:3], b[a!=b][:3])
print('bwd equal', torch.equal(g.bwd[2], ref.bwd[2]), 'rec equal', torch.equal(g.fwd_rec, ref.fwd_rec), 'nnz', g.nnz, ref.nnz, 'width', g.fwd[3], ref.fwd[3], 'M', g.M, ref.M)
