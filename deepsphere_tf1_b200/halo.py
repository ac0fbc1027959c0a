"""Halo exchange of the vertex-sharded mode (new -- the reference is single-device).

Before every sparse product each rank needs the current values of its ghost vertices (the 1-ring halo of its
contiguous vertex range).  One exchange = pack the boundary rows every neighbouring rank needs (`dsb200_gather_rows`)
and post one send / one receive per neighbour; the receives land directly in the ghost rows [Ml, Me) of the tensor,
which are grouped by owner.  torch.distributed point-to-point over NCCL (NVLink / NVSwitch) is the transport; the
same code runs over gloo on CPU tensors in the tests (with `gather=` given, no CUDA kernel).
"""
import torch
import torch.distributed as dist


def exchange(sg, t, group=None, gather=None, peer=None):
    """Fill the ghost rows of t ([1, Me, F] or [Me, F], contiguous) from their owners, in place.  `peer` (peer.PeerHalo): one
    kernel over NVLink peer memory instead of NCCL send / recv pairs; every rank must then call it for every exchange (the
    mailboxes count the calls), also ranks without neighbours."""
    if sg.world == 1:
        return t
    rows = t.view(sg.Me, -1)
    if peer is not None and peer.exchange(sg, rows):
        return t
    if not sg.send and not sg.recv:
        return t
    F = rows.shape[1]
    ops_, keep = [], []
    for q, (off, n) in sorted(sg.recv.items()):
        ops_.append(dist.P2POp(dist.irecv, rows[sg.Ml + off:sg.Ml + off + n], _peer(q, group), group))
    for q in sorted(sg.send):
        idx = sg.send_dev[q] if gather is None else sg.send[q]
        if gather is None:
            from . import ops
            buf = ops.gather_rows(rows, idx)
        else:
            buf = gather(rows, idx)
        keep.append(buf)
        ops_.append(dist.P2POp(dist.isend, buf, _peer(q, group), group))
    for w in dist.batch_isend_irecv(ops_):
        w.wait()
    return t


def _peer(q, group):
    return q if group is None else dist.get_global_rank(group, q)
