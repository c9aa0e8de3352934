"""Multi-GPU partitioning of the lnpost path: one process per GPU, grids replicated, the rows of a
batch (walkers / live points / per-star chains) sharded in contiguous blocks.  There is no data-path collective;
the only exchange is the all-gather of per-row results that a sampler's acceptance step needs.

The reference has no counterpart: its batch recipes are process pools over stars (``notebooks/batch-demo.ipynb``)
and MultiNest's own MPI (``starmodel.py:755-797``).

``RowSharder`` is pure host logic (tested with gloo, world_size 2, on CPU); ``NcclGather`` binds the library's
``ncclAllGather`` entry point on device buffers.  The 128-byte NCCL unique id travels through any byte-broadcast
the launcher offers (``torch_exchange`` for torchrun, ``file_exchange`` for a shared directory).
"""
import ctypes as C
import os
import time

import numpy as np

from . import _lib


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPU cores NVML reports as local to the GPU (its NUMA node) so that pinned host buffers
    allocated afterwards are first-touched on that node: with one process per GPU on a multi-socket box, host<->device
    DMA then stays off the inter-socket link.  Returns the CPU list, or None when NVML / affinity is unavailable."""
    try:
        import pynvml as nv

        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = device_index
        if vis:
            try:
                idx = int(vis.split(",")[device_index])
            except (ValueError, IndexError):
                idx = device_index
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        n_cpu = os.cpu_count() or 1
        words = nv.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        local = {i for i in range(n_cpu) if (int(words[i // 64]) >> (i % 64)) & 1}
        allowed = set(os.sched_getaffinity(0))
        cpus = sorted(local & allowed)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


class RowSharder(object):
    """Contiguous, balanced row blocks: rank r owns rows [start(r), stop(r)); the first ``n % world`` ranks hold one
    extra row.  Gathers are padded to ``pad`` rows per rank so that one fixed-size all-gather serves every rank."""

    def __init__(self, n_rows, world, rank=0):
        if world < 1 or not (0 <= rank < world) or n_rows < 0:
            raise ValueError("bad sharding arguments")
        self.n, self.world, self.rank = int(n_rows), int(world), int(rank)
        base, extra = divmod(self.n, self.world)
        self.counts = [base + (1 if r < extra else 0) for r in range(self.world)]
        self.starts = [sum(self.counts[:r]) for r in range(self.world)]
        self.pad = max(self.counts) if self.counts else 0

    def bounds(self, rank=None):
        r = self.rank if rank is None else rank
        return self.starts[r], self.starts[r] + self.counts[r]

    def local(self, rows, rank=None):
        a, b = self.bounds(rank)
        return rows[a:b]

    def padded(self, local_values, fill=np.nan):
        """This rank's results padded to ``pad`` entries (send buffer of the all-gather)."""
        out = np.full((self.pad,) + tuple(np.shape(local_values)[1:]), fill, dtype=np.float64)
        out[:len(local_values)] = local_values
        return out

    def assemble(self, gathered):
        """``gathered[world, pad, ...]`` (rank-major output of the all-gather) -> ``[n, ...]`` in row order."""
        g = np.asarray(gathered)
        if g.shape[:2] != (self.world, self.pad):
            raise ValueError("gathered must be [world=%d, pad=%d, ...], got %r" % (self.world, self.pad, g.shape))
        return np.concatenate([g[r, :self.counts[r]] for r in range(self.world)], axis=0)


def torch_exchange(dist):
    """Byte broadcast from rank 0 over an initialised ``torch.distributed`` group (plumbing only)."""
    def exchange(payload):
        obj = [payload]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]
    return exchange


def file_exchange(path, rank, timeout=120.0):
    """Byte broadcast from rank 0 through a file visible to every rank (no torch needed)."""
    def exchange(payload):
        if rank == 0:
            tmp = path + ".tmp"
            with open(tmp, "wb") as f:
                f.write(payload)
            os.replace(tmp, path)
            return payload
        t0 = time.time()
        while not os.path.exists(path):
            if time.time() - t0 > timeout:
                raise TimeoutError("no NCCL id at %s" % path)
            time.sleep(0.01)
        with open(path, "rb") as f:
            return f.read()
    return exchange


class FileRendezvous(object):
    """Launcher-agnostic plumbing for the ranks of ONE node through a directory every rank can see: byte all-gather,
    broadcast, barrier and a max-reduction of a float.  Each operation has a sequence number; a rank publishes
    ``<seq>.<rank>`` atomically (write + rename) and polls for the others.  Only setup and the gaps between timed
    regions go through it (handle / unique-id exchange, timing reductions) — never the data path.

    ``from_env()`` builds the directory name from the launcher's environment (``RANK`` / ``WORLD_SIZE`` as torchrun,
    mpirun wrappers or a plain shell loop export them): ``$ISO_B200_RDZV`` when set, else a name derived from the
    launching process (the ranks of one launch share their parent) and ``MASTER_PORT``."""

    def __init__(self, path, rank, world, timeout=300.0):
        self.path, self.rank, self.world, self.timeout = path, int(rank), int(world), float(timeout)
        self.seq = 0
        os.makedirs(path, exist_ok=True)

    @classmethod
    def from_env(cls, timeout=300.0):
        rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        path = os.environ.get("ISO_B200_RDZV")
        if not path:
            import tempfile

            path = os.path.join(tempfile.gettempdir(), "iso_b200_rdzv_%d_%s_%s" % (
                os.getuid(), os.environ.get("MASTER_PORT", "0"), os.getppid()))
        return cls(path, rank, world, timeout)

    def allgather_bytes(self, payload):
        """``payload -> [payload of rank 0, payload of rank 1, ...]``."""
        self.seq += 1
        mine = os.path.join(self.path, "%06d.%d" % (self.seq, self.rank))
        with open(mine + ".tmp", "wb") as f:
            f.write(payload)
        os.replace(mine + ".tmp", mine)
        out, t0 = [], time.time()
        for r in range(self.world):
            name = os.path.join(self.path, "%06d.%d" % (self.seq, r))
            while not os.path.exists(name):
                if time.time() - t0 > self.timeout:
                    raise TimeoutError("rank %d: rank %d did not reach rendezvous step %d in %s" % (self.rank, r, self.seq, self.path))
                time.sleep(0.0005)
            with open(name, "rb") as f:
                out.append(f.read())
        return out

    def broadcast(self, payload):
        """Rank 0's payload on every rank (``exchange`` of ``NcclGather``)."""
        return self.allgather_bytes(payload if self.rank == 0 else b"")[0]

    def barrier(self):
        self.allgather_bytes(b"")

    def max(self, x):
        return max(float(v) for v in self.allgather_bytes(repr(float(x)).encode()))

    def all(self, flag):
        return all(v == b"1" for v in self.allgather_bytes(b"1" if flag else b"0"))

    def close(self):
        """Every rank signs off; rank 0 removes the directory once all have (best effort)."""
        try:
            self.barrier()
            if self.rank != 0:
                open(os.path.join(self.path, "done.%d" % self.rank), "wb").close()
                return
            import shutil

            t0 = time.time()
            while not all(os.path.exists(os.path.join(self.path, "done.%d" % r)) for r in range(1, self.world)):
                if time.time() - t0 > 10.0:
                    break
                time.sleep(0.001)
            shutil.rmtree(self.path, ignore_errors=True)
        except Exception:
            pass


def preload_nccl():
    """Make ``libnccl.so.2`` resolvable for the library's ``dlopen``: the system one if present, else the copy the
    ``nvidia-nccl`` wheel ships (found through the import system, without importing torch)."""
    try:
        C.CDLL("libnccl.so.2", mode=C.RTLD_GLOBAL)
        return "libnccl.so.2"
    except OSError:
        pass
    import importlib.util

    spec = importlib.util.find_spec("nvidia.nccl")
    for base in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
        cand = os.path.join(base, "lib", "libnccl.so.2")
        if os.path.exists(cand):
            C.CDLL(cand, mode=C.RTLD_GLOBAL)
            return cand
    return None


class NcclGather(object):
    """All-gather of float64 device buffers across the ranks' contexts (``iso_nccl_init`` / ``iso_allgather_f64``).
    ``exchange(payload) -> rank 0's payload`` ships the 128-byte unique id (``FileRendezvous.broadcast``, ..)."""

    def __init__(self, ctx, rank, world, exchange):
        self.ctx, self.rank, self.world = ctx, rank, world
        preload_nccl()
        buf = C.create_string_buffer(128)
        if rank == 0:
            ctx.check(_lib.lib().iso_nccl_unique_id(buf))
        payload = exchange(bytes(buf.raw) if rank == 0 else None)
        ident = C.create_string_buffer(payload, 128)
        ctx.check(_lib.lib().iso_nccl_init(ctx.handle, ident, rank, world))

    def allgather(self, d_send, n, d_recv):
        """``d_recv[world * n] <- concat_r d_send_r[n]``; asynchronous on the context's compute stream."""
        self.ctx.check(_lib.lib().iso_allgather_f64(self.ctx.handle, d_send, int(n), d_recv))

    def close(self):
        _lib.lib().iso_nccl_destroy(self.ctx.handle)


def torch_allgather_bytes(dist):
    """``payload -> [payload of rank 0, payload of rank 1, ...]`` over an initialised ``torch.distributed`` group
    (plumbing only: ships the 128-byte peer handles)."""
    def allgather(payload):
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, payload)
        return out
    return allgather


class PeerGather(object):
    """Fused lnpost + all-gather over NVLink peer memory (``iso_peer_*``): the lnpost kernel of every rank stores its
    rows straight into the receive buffer of every rank, so the sampler's acceptance step needs no collective launch.
    One process per GPU, at most 8 ranks of one node; ``rows_per_rank`` is the (padded) block every rank contributes."""

    def __init__(self, ctx, rank, world, rows_per_rank, allgather_bytes):
        self.ctx, self.rank, self.world, self.pad = ctx, int(rank), int(world), int(rows_per_rank)
        self.handle = C.c_void_p()
        failed = None
        try:
            ctx.check(_lib.lib().iso_peer_create(ctx.handle, self.rank, self.world, self.pad, C.byref(self.handle)))
        except Exception as e:
            if self.world == 1:
                raise
            failed = e
        if self.world > 1:
            buf = C.create_string_buffer(128)
            if failed is None:
                try:
                    ctx.check(_lib.lib().iso_peer_export(ctx.handle, self.handle, buf))
                except Exception as e:
                    failed = e
            # a rank that could not create / export still joins the exchange (with zeros) before it raises
            handles = allgather_bytes(bytes(buf.raw) if failed is None else bytes(128))
            if failed is not None:
                self.close()
                raise failed
            if len(handles) != self.world or any(len(h) != 128 for h in handles):
                raise ValueError("the handle exchange must return one 128-byte payload per rank")
            blob = C.create_string_buffer(b"".join(handles), 128 * self.world)
            ctx.check(_lib.lib().iso_peer_connect(ctx.handle, self.handle, blob))

    def lnpost(self, compiled, d_pars, n, d_model_of_row=None):
        """Evaluate this rank's ``n`` rows (device buffer) and return the device pointer of the gathered
        ``[world * rows_per_rank]`` lnpost vector (rank-major); asynchronous on the context's compute stream."""
        out = C.c_void_p()
        self.ctx.check(_lib.lib().iso_lnpost_allgather_device(
            self.ctx.handle, compiled.model_pack.handle, compiled.bc_pack.handle, compiled.handle, d_model_of_row, d_pars,
            int(n), self.handle, C.byref(out)))
        return out

    def set_timeout(self, seconds):
        """Bound of the per-step completion wait (default 10 s): a rank that does not publish a step in time makes
        the next call / ``check()`` raise ``IsoError`` (``ISO_E_TIMEOUT``) instead of hanging the stream."""
        self.ctx.check(_lib.lib().iso_peer_set_timeout(self.ctx.handle, self.handle, float(seconds)))

    def check(self):
        """Synchronise the stream and raise if a step timed out."""
        self.ctx.check(_lib.lib().iso_peer_check(self.ctx.handle, self.handle))

    def close(self):
        if self.handle:
            _lib.lib().iso_peer_destroy(self.ctx.handle, self.handle)
            self.handle = C.c_void_p()


class NcclRowGather(object):
    """``PeerGather``'s interface on the library collective: the lnpost kernel writes this rank's block to a device
    buffer, ``ncclAllGather`` (on the same stream) distributes it.  What ``row_gather`` hands out when the GPUs of the
    node cannot map each other's memory (no peer access, CUDA IPC refused by the container)."""

    def __init__(self, ctx, rank, world, rows_per_rank, comm, own_comm=False):
        self.ctx, self.rank, self.world, self.pad, self.comm = ctx, int(rank), int(world), int(rows_per_rank), comm
        self._own_comm = own_comm
        self.d_own = ctx.dev_alloc(self.pad * 8)
        self.d_all = ctx.dev_alloc(self.world * self.pad * 8)
        ctx.memset(self.d_own, 0xFF, self.pad * 8)          # padding rows read as NaN

    def lnpost(self, compiled, d_pars, n, d_model_of_row=None):
        if int(n) > self.pad:
            raise ValueError("more rows than the gather was created for")
        compiled.lnpost_device(d_pars, int(n), self.d_own, d_model_of_row=d_model_of_row)
        self.comm.allgather(self.d_own, self.pad, self.d_all)
        return self.d_all

    def set_timeout(self, seconds):
        pass                     # NCCL has its own watchdog

    def check(self):
        self.ctx.sync()

    def close(self):
        if self.d_own:
            self.ctx.dev_free(self.d_own)
            self.ctx.dev_free(self.d_all)
            self.d_own = self.d_all = None
            if self._own_comm:
                self.comm.close()


def row_gather(ctx, rank, world, rows_per_rank, allgather_bytes, broadcast=None, comm=None, make_peer=None):
    """The per-step exchange of a host-driven sampler, chosen ONCE and by ALL ranks together: the fused peer-store
    gather when every rank could map every other rank's buffers, otherwise ``NcclRowGather`` (``comm``: an existing
    ``NcclGather``, or ``broadcast`` to create one).  Every rank takes part in every exchange of the decision, also a
    rank whose own setup failed — nobody waits alone.  Returns ``(gather, reason)``; ``reason`` is empty for the fused
    path and says why not otherwise."""
    make_peer = make_peer or (lambda: PeerGather(ctx, rank, world, rows_per_rank, _agreeing(allgather_bytes)))
    peer, why = None, ""
    try:
        peer = make_peer()
    except _PeerSetupFailed as e:        # some rank could not create / export its buffers: every rank lands here
        why = str(e)
    except Exception as e:               # this rank could not connect (the others may have)
        why = repr(e)[:300]
    verdicts = allgather_bytes(b"1" if peer is not None else b"0" + why.encode()[:200])
    if all(v[:1] == b"1" for v in verdicts):
        return peer, ""
    if peer is not None:
        peer.close()
    reason = why or next(v[1:].decode(errors="replace") for v in verdicts if v[:1] != b"1") or "peer setup failed on another rank"
    if comm is None:
        if broadcast is None:
            raise ValueError("row_gather: the fused path is unavailable (%s) and neither comm nor broadcast was given" % reason)
        return NcclRowGather(ctx, rank, world, rows_per_rank, NcclGather(ctx, rank, world, exchange=broadcast), own_comm=True), reason
    return NcclRowGather(ctx, rank, world, rows_per_rank, comm), reason


class _PeerSetupFailed(RuntimeError):
    pass


def _agreeing(allgather_bytes):
    """Handle exchange in which a rank that has nothing to export still takes part: ``PeerGather`` ships its 128-byte
    handle through this; a failed rank ships zeros and every rank raises together."""
    def exchange(payload):
        out = allgather_bytes(payload)
        if any(len(h) != 128 or not any(h) for h in out):
            raise _PeerSetupFailed("a rank could not export its peer buffers")
        return out
    return exchange


def sharded_lnpost(compiled, pars, sharder, gather):
    """Evaluate this rank's block of ``pars[N, ndim]`` and all-gather the results: every rank returns ``lnpost[N]``.

    ``gather(send[pad]) -> [world, pad]`` is the collective (NCCL through ``device_gather`` below on GPUs; a gloo
    all-gather in the CPU tests of the host logic)."""
    local = compiled.lnpost(np.ascontiguousarray(sharder.local(pars)))
    return sharder.assemble(gather(sharder.padded(local)))


def device_gather(ctx, comm, world):
    """Host-array front end of ``NcclGather`` (stages through device buffers; for sampler-sized batches)."""
    def gather(send):
        n = len(send)
        d_s, d_r = ctx.dev_alloc(n * 8), ctx.dev_alloc(world * n * 8)
        try:
            ctx.h2d(d_s, send)
            comm.allgather(d_s, n, d_r)
            out = np.empty((world, n))
            ctx.d2h(out, d_r)
        finally:
            ctx.dev_free(d_s)
            ctx.dev_free(d_r)
        return out
    return gather
