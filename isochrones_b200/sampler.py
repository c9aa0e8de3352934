"""On-device ensemble sampler — the GPU-side replacement for the reference's emcee driver loop
(``starmodel.py:889-972`` ``fit_mcmc_old``: ``emcee.EnsembleSampler(nwalkers, npars, self.lnpost)``, ``run_mcmc``).

``DeviceEnsembleSampler`` runs emcee's stretch move (a = 2 by default) with one persistent CTA per chain
(``iso_sampler_*`` in the C ABI): all ``n_steps`` of a run are ONE kernel launch, walkers stay in shared memory, and
``lnpost`` is the same device code the batch path uses.  ``n_chains`` independent ensembles run concurrently — with a
catalog-compiled model (one star model per chain) this is the "10k stars" mode.

Attribute names follow emcee 2.x as the reference uses them (``chain`` [walkers, steps, ndim], ``lnprobability``,
``acceptance_fraction``) for the single-chain case.
"""
import ctypes as C

import numpy as np

from . import _lib


class DeviceEnsembleSampler(object):
    def __init__(self, compiled, n_walkers, p0, seed=0, a=2.0, n_chains=1, moments=False):
        """``compiled``: ``CompiledModel`` (``BasicStarModel.compiled`` or ``compile_catalog(...)[0]``);
        ``p0``: initial walkers ``[n_walkers, ndim]`` or ``[n_chains, n_walkers, ndim]``; ``moments``: keep running
        sums of every kept (thinned) ensemble on the device (``moments()``)."""
        self.compiled = compiled
        self.ctx = compiled.ctx
        self.ndim = compiled.ndim
        p0 = np.ascontiguousarray(p0, dtype=np.float64)
        if p0.ndim == 2:
            p0 = p0[None]
        if p0.shape != (n_chains, n_walkers, self.ndim):
            raise ValueError("p0 must be [n_chains=%d, n_walkers=%d, ndim=%d], got %r"
                             % (n_chains, n_walkers, self.ndim, p0.shape))
        if compiled.n_models not in (1, n_chains):
            raise ValueError("compile one model, or one model per chain")
        self.n_chains, self.n_walkers = int(n_chains), int(n_walkers)
        self.handle = C.c_void_p()
        self.ctx.check(_lib.lib().iso_sampler_create(
            self.ctx.handle, compiled.model_pack.handle, compiled.bc_pack.handle, compiled.handle, self.n_chains,
            self.n_walkers, _lib.dp(p0), C.c_uint64(int(seed)), float(a), C.byref(self.handle)))
        self._chains = []
        self._lnprobs = []
        self.n_steps = 0
        if moments:
            self.ctx.check(_lib.lib().iso_sampler_set_moments(self.ctx.handle, self.handle, 1))

    def run_mcmc(self, n_steps, thin=1, store=True, fetch=True):
        """Advance every chain by ``n_steps`` ensemble steps (one launch).  Returns ``(pos, lnprob)`` like emcee;
        ``fetch=False`` leaves the ensemble on the device (nothing is copied to the host) and returns ``None``."""
        n_keep = n_steps // thin
        chain = np.empty((n_keep, self.n_chains, self.n_walkers, self.ndim)) if store else None
        lnp = np.empty((n_keep, self.n_chains, self.n_walkers)) if store else None
        self.ctx.check(_lib.lib().iso_sampler_run(
            self.ctx.handle, self.handle, int(n_steps), int(thin), _lib.dp(chain) if store else None,
            _lib.dp(lnp) if store else None))
        if store:
            self._chains.append(chain)
            self._lnprobs.append(lnp)
        self.n_steps += n_steps
        return self.state()[:2] if fetch else None

    def reset(self):
        """Forget the stored samples, the acceptance counters and the running moments (walker positions are kept), as
        ``emcee.EnsembleSampler.reset`` — ``acceptance_fraction`` after a burn-in + ``reset`` describes the production
        run only."""
        self._chains, self._lnprobs = [], []
        self.ctx.check(_lib.lib().iso_sampler_reset(self.ctx.handle, self.handle))

    def moments(self):
        """Posterior mean and standard deviation per chain and parameter from the running sums the kernel keeps over
        every kept (thinned) ensemble since the last ``reset``: ``(mean[n_chains, ndim], std[n_chains, ndim], count)``
        — the samples themselves never leave the GPU."""
        m = np.empty((self.n_chains, 2 * self.ndim + 1))
        self.ctx.check(_lib.lib().iso_sampler_moments(self.ctx.handle, self.handle, _lib.dp(m), None))
        cnt = m[:, -1]
        with np.errstate(invalid="ignore", divide="ignore"):
            mean = m[:, :self.ndim] / cnt[:, None]
            var = m[:, self.ndim:2 * self.ndim] / cnt[:, None] - mean ** 2
        return mean, np.sqrt(np.maximum(var, 0.0)), cnt

    def moments_device(self):
        """Device pointer of the raw running sums ``[n_chains, 2 ndim + 1]`` (send buffer of a multi-GPU gather)."""
        p = C.c_void_p()
        self.ctx.check(_lib.lib().iso_sampler_moments(self.ctx.handle, self.handle, None, C.byref(p)))
        return p

    def state(self):
        pos = np.empty((self.n_chains, self.n_walkers, self.ndim))
        lnp = np.empty((self.n_chains, self.n_walkers))
        acc = np.zeros(self.n_chains, dtype=np.int64)
        prop = C.c_int64()
        self.ctx.check(_lib.lib().iso_sampler_state(self.ctx.handle, self.handle, _lib.dp(pos), _lib.dp(lnp),
                                                    acc.ctypes.data_as(_lib.c_int64_p), C.byref(prop)))
        return pos, lnp, acc, prop.value

    # ---- stored samples ------------------------------------------------------------------------------------------
    @property
    def chains(self):
        """``[n_kept, n_chains, n_walkers, ndim]``"""
        if not self._chains:
            return np.empty((0, self.n_chains, self.n_walkers, self.ndim))
        return np.concatenate(self._chains, axis=0)

    @property
    def lnprobs(self):
        if not self._lnprobs:
            return np.empty((0, self.n_chains, self.n_walkers))
        return np.concatenate(self._lnprobs, axis=0)

    @property
    def chain(self):
        """emcee 2.x layout ``[n_walkers, n_kept, ndim]`` of chain 0 (``sampler.chain``, starmodel.py:952-969)."""
        return np.transpose(self.chains[:, 0], (1, 0, 2))

    @property
    def lnprobability(self):
        return np.transpose(self.lnprobs[:, 0], (1, 0))

    @property
    def flatchain(self):
        return self.chain.reshape(-1, self.ndim)

    @property
    def acceptance_fraction(self):
        """Per chain: accepted / proposed."""
        _, _, acc, prop = self.state()
        return acc / max(prop, 1)

    def close(self):
        if self.handle:
            _lib.lib().iso_sampler_destroy(self.ctx.handle, self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedEnsembleSampler(object):
    """ONE ensemble of walkers sharded over the GPUs of a node (``iso_ensemble_*``): every rank (one process per GPU)
    constructs this with the SAME ``p0`` and ``seed``; per half-step a rank moves only its block of the active half, and
    each proposal gathers its partner walker from the HBM of the rank that owns it over NVLink peer mappings — the
    acceptance step's exchange fused into the evaluation, no collective launch; a run ends with one
    replication of every rank's blocks, so ``state()`` reads the whole ensemble locally.  The chain is the one
    ``DeviceEnsembleSampler`` produces for the same seed, bit for bit, whatever the number of ranks; ensembles are not
    bounded by shared memory (1e5 - 1e6 walkers).

    ``allgather_bytes``: ``payload -> [payload of every rank]`` (``parallel.FileRendezvous.allgather_bytes``, a
    torch.distributed / MPI wrapper, ..); ignored for a single rank."""

    def __init__(self, compiled, n_walkers, p0, seed=0, a=2.0, rank=0, world=1, allgather_bytes=None):
        self.compiled, self.ctx, self.ndim = compiled, compiled.ctx, compiled.ndim
        self.rank, self.world, self.n_walkers = int(rank), int(world), int(n_walkers)
        p0 = np.ascontiguousarray(p0, dtype=np.float64)
        if p0.shape != (self.n_walkers, self.ndim):
            raise ValueError("p0 must be [n_walkers=%d, ndim=%d], got %r" % (self.n_walkers, self.ndim, p0.shape))
        self.handle = C.c_void_p()
        L = _lib.lib()
        self.ctx.check(L.iso_ensemble_create(
            self.ctx.handle, compiled.model_pack.handle, compiled.bc_pack.handle, compiled.handle, self.n_walkers,
            _lib.dp(p0), C.c_uint64(int(seed)), float(a), self.rank, self.world, C.byref(self.handle)))
        if self.world > 1:
            buf = C.create_string_buffer(128)
            self.ctx.check(L.iso_ensemble_export(self.ctx.handle, self.handle, buf))
            handles = allgather_bytes(bytes(buf.raw))
            if len(handles) != self.world or any(len(h) != 128 for h in handles):
                raise ValueError("the handle exchange must return one 128-byte payload per rank")
            blob = C.create_string_buffer(b"".join(handles), 128 * self.world)
            self.ctx.check(L.iso_ensemble_connect(self.ctx.handle, self.handle, blob))
        self._chains, self._lnprobs = [], []
        self.n_steps = 0

    def set_timeout(self, seconds):
        self.ctx.check(_lib.lib().iso_ensemble_set_timeout(self.ctx.handle, self.handle, float(seconds)))

    def run_mcmc(self, n_steps, thin=1, store=True, fetch=True):
        """Advance the ensemble by ``n_steps`` steps (every rank calls this with the same arguments); ``store`` keeps
        the thinned chain on THIS rank.  Returns ``(pos[n_walkers, ndim], lnprob[n_walkers])``, or ``None`` with
        ``fetch=False`` (the ensemble stays on the device)."""
        n_keep = n_steps // thin
        chain = np.empty((n_keep, self.n_walkers, self.ndim)) if store else None
        lnp = np.empty((n_keep, self.n_walkers)) if store else None
        self.ctx.check(_lib.lib().iso_ensemble_run(self.ctx.handle, self.handle, int(n_steps), int(thin),
                                                   _lib.dp(chain) if store else None, _lib.dp(lnp) if store else None))
        if store:
            self._chains.append(chain)
            self._lnprobs.append(lnp)
        self.n_steps += n_steps
        return self.state()[:2] if fetch else None

    def state(self):
        """``(pos, lnprob, proposals accepted by this rank, proposals per ensemble)``."""
        pos, lnp = np.empty((self.n_walkers, self.ndim)), np.empty(self.n_walkers)
        acc, prop = C.c_int64(), C.c_int64()
        self.ctx.check(_lib.lib().iso_ensemble_state(self.ctx.handle, self.handle, _lib.dp(pos), _lib.dp(lnp), C.byref(acc),
                                                     C.byref(prop)))
        return pos, lnp, acc.value, prop.value

    @property
    def chains(self):
        """``[n_kept, n_walkers, ndim]``"""
        return np.concatenate(self._chains, axis=0) if self._chains else np.empty((0, self.n_walkers, self.ndim))

    @property
    def lnprobs(self):
        return np.concatenate(self._lnprobs, axis=0) if self._lnprobs else np.empty((0, self.n_walkers))

    def close(self):
        if self.handle:
            _lib.lib().iso_ensemble_destroy(self.ctx.handle, self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
