"""x-slab decomposition of one lattice over several GPUs.

x is the slowest index of the reference layout, so the slab [x0, x1) of every population array is a
contiguous range; x is periodic in every multiphase case, so the slabs form a ring (rank R-1 <-> rank 0).
One time step of a slab is

    stage 0 : moments of the local planes, pack the moment halo          (clbm_step_stage(0))
    exchange phase 0 with both neighbours
    stage 1 : unpack, collide + push-stream, pack the crossing populations (clbm_step_stage(1))
    exchange phase 1
    stage 2 : unpack the crossing populations into the boundary planes     (clbm_step_stage(2))

The exchange itself is plumbing.  On GPUs the default is the library's own peer-memory ring (csrc/slab_comm.cu: the pack
writes straight into the neighbour's mailbox over NVLink, flag signal / wait kernels, two steps per CUDA-graph launch;
`clbm_slab_step`), for which this module only moves the 128-byte mailbox handles once.  `LocalRing` runs several contexts
in one process (single-GPU emulation of R ranks, used by the GPU tests; copies between the buffers, or the same peer
ring), `DistRing` is one rank per process; its "torch" transport moves the buffers with torch.distributed point-to-point
operations (NCCL on GPUs; gloo in the CPU tests with host buffers).
"""
import os

import numpy as np


def slab_bounds(nx_global, nranks):
    """[x0, x1) of every rank; the remainder goes to the first ranks."""
    base, rem = divmod(nx_global, nranks)
    out, x = [], 0
    for r in range(nranks):
        w = base + (1 if r < rem else 0)
        out.append((x, x + w))
        x += w
    return out


def slab_params(params, rank, nranks):
    """clbm_params of rank's slab of a global lattice described by `params` (nx == nx_global)."""
    x0, x1 = slab_bounds(params.nx_global, nranks)[rank]
    return params.copy(nx=x1 - x0, x_offset=x0)


def slice_host_state(params, lattice, flag, rank, nranks):
    """cut the reference-layout host arrays of the global lattice down to one slab (same layout)."""
    x0, x1 = slab_bounds(params.nx_global, nranks)[rank]
    plane = params.ny * params.nz
    ne_g = params.nx_global * plane
    Q, sets = params.Q, params.sets
    lat = lattice.reshape(sets, 2, Q, ne_g)[:, :, :, x0 * plane:x1 * plane]
    return np.ascontiguousarray(lat).reshape(-1), np.ascontiguousarray(flag[x0 * plane:x1 * plane])


class CudaBuffer:
    """view of a raw device pointer for torch (CUDA array interface)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def as_torch(ptr, nbytes, device):
    import torch
    return torch.as_tensor(CudaBuffer(ptr, nbytes), device=device)


def combine_contact_angle(ny, first, second):
    """the two rounds of clbm_diag_contact_angle_slab of all slabs -> (base_y, base, height) of calculateContactAngle
    (SC/apps/contactAngle2D.h:465-505); `second` is None when the first round found no fluid row"""
    base_y = min(r[0] for r in first)
    if base_y >= ny - 1 or second is None:
        return base_y, 0, 0
    lstop, rstop, hstop = max(r[1] for r in second), min(r[2] for r in second), min(r[3] for r in second)
    return base_y, max(0, rstop - lstop - 1), hstop - base_y


class LocalRing:
    """R slab contexts in one process; the exchange is a device-to-device copy per neighbour pair."""

    def __init__(self, lattices, peer=False):
        """peer=True: the library's own peer-memory ring (clbm_peer_connect_local + clbm_slab_step: packs straight into the
        neighbour's mailbox, flag signal / wait kernels, CUDA-graph replay) instead of the copies below"""
        import torch
        self.lats = lattices
        self.torch = torch
        self.R = len(lattices)
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self._views = {}
        self.peer = bool(peer) and self.R > 1
        if self.peer:
            for r, lat in enumerate(lattices):
                lat.peer_connect_local(lattices[(r - 1) % self.R], lattices[(r + 1) % self.R])

    def _view(self, r, phase, side, recv):
        key = (r, phase, side, recv)
        if key not in self._views:
            ptr, nb = self.lats[r].halo_buffer(phase, side, recv)
            self._views[key] = as_torch(ptr, nb, self.dev)
        return self._views[key]

    def exchange(self, phase):
        for lat in self.lats:
            lat.sync()
        for r in range(self.R):
            left, right = (r - 1) % self.R, (r + 1) % self.R
            # my side-0 send buffer travels to the left neighbour's side-1 recv buffer, and vice versa
            self._view(left, phase, 1, True).copy_(self._view(r, phase, 0, False))
            self._view(right, phase, 0, True).copy_(self._view(r, phase, 1, False))
        self.torch.cuda.synchronize()

    def _peer_phase(self, phase, boundary):
        """one exchange of the peer ring for all contexts: EVERY signal before ANY wait.  The contexts share one GPU here, and
        a spinning wait kernel of one context may sit in front of another context's signal in a hardware queue the two
        streams happen to share; with all signals submitted first no wait can starve (one context per GPU has no such issue:
        clbm_slab_step then runs whole steps, two per CUDA-graph launch)."""
        for lat in self.lats:
            lat.slab_signal(phase, boundary)
        for lat in self.lats:
            lat.slab_wait(phase, boundary)

    def exchange_flags(self):
        if self.peer:
            for lat in self.lats:
                lat.halo_pack(2)
            self._peer_phase(2, False)
            for lat in self.lats:
                lat.halo_unpack(2)
                lat.sync()
            return
        for lat in self.lats:
            lat.halo_pack(2)
        self.exchange(2)
        for lat in self.lats:
            lat.halo_unpack(2)
            lat.sync()

    def step(self, n=1, overlap=None):
        """overlap: use the boundary-first protocol (stages 10-12) where the contexts support it.  Copy ring: the exchange is a
        synchronous copy, so this only checks the protocol's results, not its timing.  Peer ring: the stages of all contexts
        are interleaved with the signal / wait kernels (no host synchronisation anywhere; see _peer_phase for the order)."""
        if self.peer:
            if overlap is None:
                overlap = all(lat.overlap_supported() for lat in self.lats)
            s0, s1, s2 = (10, 11, 12) if overlap else (0, 1, 2)
            for _ in range(n):
                for lat in self.lats:
                    lat.step_stage(s0)
                self._peer_phase(0, overlap and self.lats[0].overlap_variant() == 1)   # form 1 moves the moment halo on the boundary stream
                for lat in self.lats:
                    lat.step_stage(s1)
                self._peer_phase(1, overlap)
                for lat in self.lats:
                    lat.step_stage(s2)
            return
        if overlap is None:
            overlap = all(lat.overlap_supported() for lat in self.lats)
        s0, s1, s2 = (10, 11, 12) if overlap else (0, 1, 2)
        for _ in range(n):
            for lat in self.lats:
                lat.step_stage(s0)
            self.exchange(0)
            for lat in self.lats:
                lat.step_stage(s1)
            self.exchange(1)
            for lat in self.lats:
                lat.step_stage(s2)

    def contact_angle_scan(self, rho_cut):
        """calculateContactAngle's scans over all slabs (two rounds per slab, combined here)"""
        ny = self.lats[0].p.ny
        first = [lat.contact_angle_scan_slab(rho_cut, -1) for lat in self.lats]
        base_y = min(r[0] for r in first)
        second = [lat.contact_angle_scan_slab(rho_cut, base_y) for lat in self.lats] if base_y < ny - 1 else None
        return combine_contact_angle(ny, first, second)

    def interface_heights(self, phi_mid):
        """findInterfaceHeights over all slabs: a slab answers 0 for a column it does not own"""
        r = [lat.interface_heights(phi_mid) for lat in self.lats]
        return max(v[0] for v in r), max(v[1] for v in r)

    def refresh_moment_halo(self):
        """make the moments and their ghosts valid for the current populations (needed before fields()); stage 20 = stage 0
        with every moment rebuilt from the populations (the sweep kernel of HCZ D3Q19 keeps its moments split in partial sums)"""
        if self.peer:
            for lat in self.lats:
                lat.step_stage(20)
            self._peer_phase(0, False)
            for lat in self.lats:
                lat.halo_unpack(0)
                lat.sync()
            return
        for lat in self.lats:
            lat.step_stage(20)
        self.exchange(0)
        for lat in self.lats:
            lat.halo_unpack(0)
            lat.sync()


def ring_exchange(dist, rank, nranks, send0, send1, recv0, recv1, group=None):
    """One exchange phase with both ring neighbours through torch.distributed P2P ops.
    send0 goes to rank-1 (arriving in its recv1), send1 to rank+1 (arriving in its recv0)."""
    if nranks == 1:
        recv1.copy_(send0)
        recv0.copy_(send1)
        return
    left, right = (rank - 1) % nranks, (rank + 1) % nranks
    ops = [dist.P2POp(dist.isend, send0, left, group), dist.P2POp(dist.isend, send1, right, group),
           dist.P2POp(dist.irecv, recv1, right, group), dist.P2POp(dist.irecv, recv0, left, group)]
    if nranks == 2:
        # left == right: tag-less P2P matches in posting order; order the ops identically on both ranks
        ops = [dist.P2POp(dist.isend, send0, left, group), dist.P2POp(dist.irecv, recv1, right, group),
               dist.P2POp(dist.isend, send1, right, group), dist.P2POp(dist.irecv, recv0, left, group)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()


class DistRing:
    """one slab per process (rank), exchange over torch.distributed (NCCL send/recv over NVLink).

    The library's launching stream is made torch's current stream (torch.cuda.ExternalStream) for the duration of
    every ring call, so the NCCL operations are ordered after the pack kernels and before the unpack kernels on the
    DEVICE: `req.wait()` on an NCCL work object makes the current stream wait, not the host.  A slab step therefore
    never synchronises with the host; the step loop runs ahead of the GPU like the single-slab one."""

    def __init__(self, lattice, rank, nranks, device, native=None, transport=None):
        import torch
        import torch.distributed as dist
        self.lat, self.rank, self.R, self.dist, self.torch = lattice, rank, nranks, dist, torch
        # native ring (csrc/slab_comm.cu): the step loop, the stages and the ncclSend/ncclRecv groups all run inside the
        # library; torch.distributed only carries the 128-byte ncclUniqueId once.  Opt-in (CLBM_SLAB_NATIVE=1) until it has
        # been measured on a multi-GPU box: it exists because a slab step of BASELINE configs[2] (256 columns per GPU, a
        # 0.17 ms kernel) is host-launch bound when the five calls per step are issued from Python.
        if native is None:
            native = os.environ.get("CLBM_SLAB_NATIVE", "0") == "1"
        self.native = bool(native) and device.type == "cuda" and nranks > 1
        self.dev = device
        # transport of the ghost exchange: "peer" (default on GPUs: the library's peer-memory ring over CUDA IPC, packs
        # straight into the neighbour's mailbox, CUDA-graph replay of the slab step), "nccl" (library-driven ncclSend/Recv),
        # "torch" (torch.distributed batch_isend_irecv issued from Python, the round-1 path; the only one on CPU / gloo)
        want = transport or os.environ.get("CLBM_SLAB_TRANSPORT", "") or ("nccl" if self.native else "peer")
        if device.type != "cuda" or nranks < 2:
            want = "torch"
        self.transport = want
        if want == "peer":
            try:
                self._init_peer()
            except Exception as e:   # noqa: BLE001  (IPC not permitted in this container, ...): say so and fall back
                import sys
                print("clbm: peer-memory ring unavailable (%s); falling back to torch.distributed P2P" % e, file=sys.stderr)
                self.transport = "torch"
                ok = 0
            else:
                ok = 1
            # every rank must agree on the transport
            t = torch.tensor([ok], dtype=torch.int32, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if int(t.item()) == 0 and self.transport == "peer":
                lattice.peer_disconnect()
                self.transport = "torch"
        self.native = self.transport == "nccl"
        self._v = {}
        for phase in range(3):
            for side in range(2):
                for recv in (False, True):
                    ptr, nb = lattice.halo_buffer(phase, side, recv)
                    self._v[(phase, side, recv)] = as_torch(ptr, nb, device)
        self.stream = torch.cuda.ExternalStream(lattice.stream(), device=device) if device.type == "cuda" else None
        # boundary-first overlap protocol (clbm_step_stage 10-12): the exchanges run on the library's boundary stream
        self.overlap = device.type == "cuda" and lattice.overlap_supported()
        self.stream_b = torch.cuda.ExternalStream(lattice.boundary_stream(), device=device) if self.overlap else None
        # The NCCL kernels run on the process group's own stream.  With the default (normal-priority) stream they queue
        # behind the thousands of pending CTAs of the interior launch and the "overlapped" exchange only starts when the
        # interior kernel drains; a group whose NCCL streams are high priority gets the next SM slot that frees up.
        self.group = None
        if self.overlap and nranks > 1 and self.transport == "torch":
            try:
                opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
                self.group = dist.new_group(list(range(nranks)), pg_options=opts)
            except Exception:   # noqa: BLE001  (gloo / older torch: keep the default group)
                self.group = None

    def _on_stream(self):
        import contextlib
        return self.torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()

    def exchange(self, phase, boundary=False):
        v = self._v
        with (self.torch.cuda.stream(self.stream_b) if boundary else self._on_stream()):
            ring_exchange(self.dist, self.rank, self.R, v[(phase, 0, False)], v[(phase, 1, False)],
                          v[(phase, 0, True)], v[(phase, 1, True)], group=self.group if boundary else None)

    def record_event(self):
        """CUDA event on the launching stream (device-side timing of slab steps)"""
        ev = self.torch.cuda.Event(enable_timing=True)
        ev.record(self.stream)
        return ev

    def exchange_flags(self):
        if self.transport == "peer":
            self.lat.slab_exchange(2)
            self.lat.sync()
            return
        self.lat.halo_pack(2)
        self.exchange(2)
        self.lat.halo_unpack(2)
        self.lat.sync()

    def contact_angle_scan(self, rho_cut):
        """calculateContactAngle's scans over the ring: two rounds on every rank, MIN / MAX all-reduces in between"""
        torch, dist = self.torch, self.dist
        ny = self.lat.p.ny
        first = self.lat.contact_angle_scan_slab(rho_cut, -1)
        t = torch.tensor([first[0]], dtype=torch.int64, device=self.dev)
        if self.R > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        base_y = int(t.item())
        if base_y >= ny - 1:
            return base_y, 0, 0
        r = self.lat.contact_angle_scan_slab(rho_cut, base_y)
        t = torch.tensor([r[1], -r[2], -r[3]], dtype=torch.int64, device=self.dev)      # one MAX for max / min / min
        if self.R > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        lstop, rstop, hstop = int(t[0]), -int(t[1]), -int(t[2])
        return base_y, max(0, rstop - lstop - 1), hstop - base_y

    def interface_heights(self, phi_mid):
        torch, dist = self.torch, self.dist
        t = torch.tensor(list(self.lat.interface_heights(phi_mid)), dtype=torch.int64, device=self.dev)
        if self.R > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return int(t[0]), int(t[1])

    def _init_peer(self):
        """all-gather the 128-byte mailbox handles, map both ring neighbours (clbm_peer_connect), barrier"""
        torch, dist = self.torch, self.dist
        mine = torch.frombuffer(bytearray(self.lat.peer_export()), dtype=torch.uint8).to(self.dev)
        every = [torch.empty_like(mine) for _ in range(self.R)]
        dist.all_gather(every, mine)
        left, right = (self.rank - 1) % self.R, (self.rank + 1) % self.R
        self.lat.peer_connect(every[left].cpu().numpy().tobytes(), every[right].cpu().numpy().tobytes())
        dist.barrier()

    def _init_native(self):
        torch, dist = self.torch, self.dist
        t = torch.zeros(128, dtype=torch.uint8, device=self.dev)
        if self.rank == 0:
            t.copy_(torch.frombuffer(bytearray(self.lat.comm_unique_id()), dtype=torch.uint8))
        if self.R > 1:
            dist.broadcast(t, 0)
        self.lat.comm_init(bytes(t.cpu().numpy().tobytes()), self.rank, self.R)
        self._native_ready = True

    def step(self, n=1, overlap=None):
        if self.transport == "peer":
            self.lat.slab_step(n)
            return
        if self.native and overlap is None:
            if not getattr(self, "_native_ready", False):
                self._init_native()
            self.lat.slab_step(n)
            return
        if self.overlap if overlap is None else (overlap and self.overlap):
            for _ in range(n):
                self.lat.step_stage(10)         # boundary moments + pack (launching stream)
                self.exchange(0, boundary=self.lat.overlap_variant() == 1)
                self.lat.step_stage(11)         # boundary planes (form 2: then the interior) | boundary stream: pack of the crossing populations
                self.exchange(1, boundary=True)
                self.lat.step_stage(12)         # unpack, join
            return
        with self._on_stream():
            for _ in range(n):
                self.lat.step_stage(0)
                self.exchange(0)
                self.lat.step_stage(1)
                self.exchange(1)
                self.lat.step_stage(2)

    def refresh_moment_halo(self):
        if self.transport == "peer":
            self.lat.step_stage(20)
            self.lat.slab_exchange(0)
            self.lat.sync()
            return
        self.lat.step_stage(20)
        self.exchange(0)
        self.lat.halo_unpack(0)
        self.lat.sync()
