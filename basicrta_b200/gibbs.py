"""Drop-in ``Gibbs`` / ``ParallelGibbs`` with the sweep on a B200.

Mirrors the public surface of the reference's ``basicrta/gibbs.py`` for the hot path:

* ``Gibbs(times, residue, loc, ncomp, niter, cutoff)`` with the same attributes
  (gibbs.py:133-157), ``run()`` filling ``mcweights`` / ``mcrates`` / ``indicator`` thinned
  every ``g`` steps and pickling the instance to ``basicrta-{cutoff}/{residue}/
  gibbs_{niter}.pkl`` (gibbs.py:176-219, 336-349), ``load`` (gibbs.py:351-381),
  ``__getitem__`` (gibbs.py:159-160), ``_sample_indicator`` (gibbs.py:321-334);
* ``ParallelGibbs(contacts, nproc, ncomp, niter).run(run_resids)`` (gibbs.py:20-88):
  residues are sharded over ``nproc`` GPUs instead of a multiprocessing pool.

``g``, ``burnin``, ``whypers``, ``rhypers`` are read at ``run()`` time, so the reference's
"assign after construction" idiom (tests/test_functions.py:11-12) keeps working.

The sweep itself (gibbs.py:191-217) is ``brta_gibbs_run_batch`` of the C ABI; there is no
CPU implementation in this package.  Posterior processing (``cluster``, ``process_gibbs``,
``estimate_tau``) stays on the host, plot-free, in :mod:`basicrta_b200.postprocess`.

Output path: the labels of a batch come back in chunks through pinned staging
buffers while worker threads scatter them into the per-residue arrays and write the pickles; the
pickle stream names the reference's classes (``basicrta.gibbs.Gibbs``,
``MDAnalysis.analysis.base.Results``), so a stock basicrta loads it without this package.
"""
import os
import pickle
import struct
import threading
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _cabi
from .engine import ChainInput, coefficients, get_engine, tick_grid
from .plan import shard_chains
from .util import get_s

try:                                                   # same class as the reference when available
    from MDAnalysis.analysis.base import Results       # gibbs.py:9
except Exception:                                      # MDAnalysis is optional here
    from .results import Results


def _fresh_seed():
    return int(np.random.SeedSequence().entropy) & 0xFFFFFFFFFFFFFFFF


def _chain_id(residue, cutoff):
    return zlib.crc32(f'{residue}|{cutoff}'.encode()) & 0xFFFFFFFF


class GibbsBatchError(RuntimeError):
    """Some residues of a batch failed; every other residue was sampled (and saved) normally.
    ``failures``: list of (residue, exception)."""

    def __init__(self, failures):
        self.failures = list(failures)
        lines = '; '.join(f'{res}: {type(e).__name__}: {e}' for res, e in self.failures[:8])
        more = '' if len(self.failures) <= 8 else f' ... and {len(self.failures) - 8} more'
        super().__init__(f'{len(self.failures)} residue(s) failed -- {lines}{more}')


# attributes that only exist at run time and never enter the pickle
_TRANSIENT = ('_device_indicator', '_device_batch')


class Gibbs(object):
    """Gibbs sampler for an exponential mixture of residence times (one residue).

    Same constructor and attributes as ``basicrta.gibbs.Gibbs`` (gibbs.py:133-157).  Two
    attributes are new and optional: ``seed`` (None = fresh entropy per run; the reference
    never seeds, gibbs.py:17) and ``device`` (CUDA ordinal; default ``loc % device_count``).
    """

    def __init__(self, times=None, residue=None, loc=0, ncomp=15, niter=110000, cutoff=None):
        self.times = times
        self.residue = residue
        self.niter = niter
        self.loc = loc
        self.ncomp = ncomp
        self.g = 100
        self.burnin = 10000
        self.cutoff = cutoff
        self.processed_results = Results()
        self._noise_cutoff = 0.4

        if times is not None:
            srt = np.sort(times)
            gaps = srt[1:] - srt[:-1]
            nonzero = gaps[gaps != 0]
            self.ts = nonzero[0] if len(nonzero) else times.min()
        else:
            self.ts = None

        self.keys = {'times', 'residue', 'loc', 'ncomp', 'niter', 'g', 'burnin',
                     'processed_results', 'ts', 'mcweights', 'mcrates', 't',
                     's', 'cutoff', 'indicator'}

    def __getitem__(self, item):
        return getattr(self, item)

    def __getstate__(self):
        return {k: v for k, v in self.__dict__.items() if k not in _TRANSIENT}

    # ---- host preparation (gibbs.py:162-174) ---------------------------------------------
    def _prepare(self, allocate_indicator=True):
        self.t, self.s = get_s(self.times, self.ts)
        rows = (self.niter + 1) // self.g
        # the batch path fills `indicator` from the device and skips this zero-filled allocation
        self.indicator = np.zeros((rows, self.times.shape[0]), dtype=np.uint8) if allocate_indicator else None
        self.mcweights = np.zeros((rows, self.ncomp))
        self.mcrates = np.zeros((rows, self.ncomp))
        self.whypers = np.ones(self.ncomp) / [self.ncomp]
        self.rhypers = np.ones((self.ncomp, 2)) * [1, 3]

    def _chain_input(self):
        """The chain in device units.  The device grid is NOT ``self.ts`` (the reference's first gap,
        kept for the pickle) but the coarsest grid all times lie on, see :func:`engine.tick_grid`."""
        ticks, grid = tick_grid(self.times, self.ts)
        return ChainInput(ticks=ticks, ts=float(grid),
                          chain_id=_chain_id(self.residue, self.cutoff),
                          whypers=np.asarray(self.whypers, dtype=np.float64),
                          rhypers=np.asarray(self.rhypers, dtype=np.float64))

    def _savedir(self):
        return f'basicrta-{self.cutoff}/{self.residue}'

    # ---- the sampler (gibbs.py:176-219) ---------------------------------------------------
    def run(self, keep_indicator_on_device=False):
        """Run the sampler on the GPU and pickle the instance, like the reference's ``run``.

        ``keep_indicator_on_device``: the label rows stay in HBM (``self.indicator`` is None, which the
        reference's ``cluster`` understands, gibbs.py:259-262); ``process_gibbs(device=...)`` /
        ``cluster(device=...)`` then reduce them to cluster counts on the GPU and only ``[N, clusters]``
        integers cross PCIe."""
        self._prepare(allocate_indicator=False)
        os.makedirs(self._savedir(), exist_ok=True)
        try:
            run_batch([self], device=getattr(self, 'device', None), prepared=True, save=True,
                      keep_indicator_on_device=keep_indicator_on_device)
        except GibbsBatchError as e:
            raise e.failures[0][1] from None

    def _sample_indicator(self):
        """Re-draw the labels of every stored (mcweights, mcrates) row without a parameter
        update (gibbs.py:321-334): the sweep kernel in teacher-forced mode, thin = 1."""
        rows = self.mcweights.shape[0]
        ticks, grid = tick_grid(self.times, self.ts)
        coef = [coefficients(w, r, grid) for w, r in zip(self.mcweights, self.mcrates)]
        eng = get_engine(_pick_device(getattr(self, 'device', None), self.loc))
        chain = ChainInput(ticks=ticks, ts=float(grid), chain_id=_chain_id(self.residue, self.cutoff))
        seed = getattr(self, 'seed', None)
        res = eng.run([chain], self.ncomp, rows, thin=1, seed=_fresh_seed() if seed is None else seed,
                      flags=_cabi.FLAG_INJECT_COEF,
                      inject={'coef_c': [np.stack([c for c, _ in coef])],
                              'coef_a': [np.stack([a for _, a in coef])]})[0]
        err = _status_error(self.residue, res.status)
        if err is not None:
            raise err
        self.indicator = np.ascontiguousarray(res.indicator[:rows])   # (niter+1)//1 rows are allocated, `rows` filled
        return self.indicator[self.burnin // self.g:]

    def device_indicator(self):
        """The label rows still resident on the GPU (``run(keep_indicator_on_device=True)``), as a CUDA
        uint8 tensor ``[rows, N]``, or None."""
        return getattr(self, '_device_indicator', None)

    # ---- persistence (gibbs.py:336-381) ---------------------------------------------------
    def save(self):
        savedir = self._savedir() + '/'
        filename = f'gibbs_{self.niter}.pkl'
        if not os.path.exists(savedir):
            raise OSError(f'No such directory: {savedir}')
        if os.path.exists(savedir + filename):
            os.rename(savedir + filename, savedir + filename + '.bak')
        with open(savedir + filename, 'w+b') as f:
            dump_reference_pickle(self, f)

    def save_compact(self, path=None):
        """Opt-in compact side-car of a finished chain: everything ``save`` writes, with the
        label rows bit-packed (``util.pack_labels``: 4 bits per label for ncomp <= 16 -- half the bytes of the
        reference pickle, a quarter for ncomp <= 4) in an uncompressed ``.npz``.  The reference pickle stays the
        default and the only format stock basicrta reads; :meth:`load_compact` restores the same object."""
        from .util import pack_labels
        if self.indicator is None:
            raise ValueError('no label rows on the host (run with keep_indicator_on_device?)')
        path = path or f'{self._savedir()}/gibbs_{self.niter}.compact.npz'
        import json
        scalars = dict(residue=str(self.residue), loc=int(self.loc), ncomp=int(self.ncomp), niter=int(self.niter),
                       g=int(self.g), burnin=int(self.burnin), noise_cutoff=float(self._noise_cutoff),
                       cutoff=None if self.cutoff is None else float(self.cutoff))
        with open(path, 'wb') as f:
            np.savez(f, packed=pack_labels(self.indicator, self.ncomp), n_data=len(self.times), times=self.times,
                     mcweights=self.mcweights, mcrates=self.mcrates, t=self.t, s=self.s, whypers=self.whypers,
                     rhypers=self.rhypers, scalars=np.array(json.dumps(scalars)))
        return path

    @staticmethod
    def load_compact(file):
        """Restore a chain written by :meth:`save_compact` (labels unpacked to the reference's uint8 rows)."""
        import json
        from .util import unpack_labels
        with np.load(file) as z:
            sc = json.loads(str(z['scalars']))
            g = Gibbs(z['times'], sc['residue'], sc['loc'], ncomp=sc['ncomp'], niter=sc['niter'], cutoff=sc['cutoff'])
            g.g, g.burnin, g._noise_cutoff = sc['g'], sc['burnin'], sc['noise_cutoff']
            g.mcweights, g.mcrates, g.t, g.s = z['mcweights'], z['mcrates'], z['t'], z['s']
            g.whypers, g.rhypers = z['whypers'], z['rhypers']
            g.indicator = unpack_labels(z['packed'], g.ncomp, int(z['n_data']))
        return g

    @staticmethod
    def load(file):
        keys = ['times', 'residue', 'loc', 'ncomp', 'niter', 'g', 'burnin',
                'processed_results', 'ts', 'mcweights', 'mcrates', 't',
                's', 'cutoff', 'indicator', 'whypers', 'rhypers']
        with open(file, 'r+b') as f:
            r = load_reference_pickle(f)
        g = Gibbs()
        for attr in keys:
            try:
                setattr(g, attr, r[f'{attr}'])
            except AttributeError:
                setattr(g, attr, None)
        if isinstance(g.residue, np.ndarray):
            g.residue = g.residue[0]
        if g.t is None:
            g.t, g.s = get_s(g.times, g.ts)
        return g

    # ---- posterior processing: host side, see postprocess.py -------------------------------
    def cluster(self, method='GaussianMixture', device=None, gmm_device=None, **kwargs):
        """gibbs.py:221-273.  ``device``: GPU for the membership counts; ``gmm_device``: GPU for the mixture fit
        (``basicrta_b200.gmm``) instead of scikit-learn."""
        from . import postprocess
        return postprocess.cluster(self, method=method, device=device, gmm_device=gmm_device, **kwargs)

    def process_gibbs(self, save=True, device=None, gmm_device=None):
        from . import postprocess
        return postprocess.process_gibbs(self, save=save, device=device, gmm_device=gmm_device)

    def estimate_tau(self):
        from . import postprocess
        return postprocess.estimate_tau(self)


# ---- pickles that a stock basicrta can read -------------------------------------------------------
# The reference pickles the whole instance (gibbs.py:347) and its loader needs the class importable as
# ``basicrta.gibbs.Gibbs`` and ``processed_results`` as ``MDAnalysis.analysis.base.Results``
# (gibbs.py:9, 351-381).  The stream written here names exactly those, whatever this package is called,
# and writes the big arrays straight from their memory (protocol 5, in band): no intermediate copy, and
# the file write releases the GIL, so a pool of writer threads scales.
def _class_aliases():
    from .results import Results as OwnResults
    return {Gibbs: ('basicrta.gibbs', 'Gibbs'), OwnResults: ('MDAnalysis.analysis.base', 'Results')}


class _RefPickler(pickle._Pickler):
    dispatch = pickle._Pickler.dispatch.copy()

    def __init__(self, file):
        super().__init__(file, protocol=5)
        self._aliases = _class_aliases()

    def save_global(self, obj, name=None):
        alias = self._aliases.get(obj) if isinstance(obj, type) else None
        if alias is None:
            return super().save_global(obj, name)
        self.save(alias[0])
        self.save(alias[1])
        self.write(pickle.STACK_GLOBAL)
        self.memoize(obj)

    def _save_buffer_in_band(self, obj):
        # a buffer is written once and nothing else refers to it, so it needs no memo entry
        with obj.raw() as m:
            header = (pickle.BINBYTES8 if m.readonly else pickle.BYTEARRAY8) + struct.pack('<Q', m.nbytes)
            if m.nbytes >= self.framer._FRAME_SIZE_TARGET:
                self._write_large_bytes(header, m)          # straight from the array's memory
            else:
                self.write(header + m.tobytes())

    dispatch[pickle.PickleBuffer] = _save_buffer_in_band


class _RefUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) == ('basicrta.gibbs', 'Gibbs'):
            return Gibbs
        if (module, name) == ('MDAnalysis.analysis.base', 'Results'):
            return Results
        return super().find_class(module, name)


class _DeferredPickler(_RefPickler):
    """The reference pickle with HOLES: the payload of selected arrays is not written; the stream records where it
    belongs (``offsets[name]`` = byte position in the file) and skips that many bytes.  Everything else -- opcodes,
    memo indices, the length prefix of the skipped payload -- is the stream ``_RefPickler`` writes (up to the
    placement of FRAME marks, which unpicklers skip), so filling the holes later yields the pickle of the
    finished object."""
    dispatch = _RefPickler.dispatch.copy()

    def __init__(self, file, deferred):
        super().__init__(file)
        self._out = file
        self._deferred = {int(a.__array_interface__['data'][0]): name for name, a in deferred.items()}
        self.offsets = {}

    def _save_buffer_deferred(self, obj):
        with obj.raw() as m:
            name = self._deferred.get(int(np.asarray(m).__array_interface__['data'][0])) if m.nbytes else None
            if name is None:
                return _RefPickler._save_buffer_in_band(self, obj)
            header = (pickle.BINBYTES8 if m.readonly else pickle.BYTEARRAY8) + struct.pack('<Q', m.nbytes)
            self.framer.commit_frame(force=True)               # everything so far is in the file
            self._out.write(header)
            self.offsets[name] = self._out.tell()
            self._out.seek(m.nbytes, os.SEEK_CUR)              # the hole

    dispatch[pickle.PickleBuffer] = _save_buffer_deferred


class DeferredPickle:
    """``gibbs_{niter}.pkl`` of a residue created BEFORE its chain has finished: the label array of the file is
    a hole that the output path fills in place (``indicator`` is a memory map of that region, so bringing the
    labels home IS writing the file), ``mcweights`` / ``mcrates`` are patched in at the end.  Same opcodes and
    payloads as ``Gibbs.save`` of the finished object, same ``.bak`` rotation (gibbs.py:336-349)."""

    def __init__(self, gb, rows):
        self.gb = gb
        savedir = gb._savedir() + '/'
        self.path = savedir + f'gibbs_{gb.niter}.pkl'
        if not os.path.exists(savedir):
            raise OSError(f'No such directory: {savedir}')
        self.rotated = os.path.exists(self.path)
        if self.rotated:
            os.rename(self.path, self.path + '.bak')
        n = len(gb.times)
        gb.indicator = np.empty((rows, n), dtype=np.uint8)     # never touched: stands for the hole
        with open(self.path, 'w+b') as f:
            pk = _DeferredPickler(f, {'indicator': gb.indicator, 'mcweights': gb.mcweights, 'mcrates': gb.mcrates})
            pk.dump(gb)
        self.offsets = pk.offsets
        self.shape = (rows, n)
        self._map = None
        self.fd = os.open(self.path, os.O_RDWR)                # the staging threads pwrite the label blocks through it
        gb.indicator = None

    @property
    def sink(self):
        """(fd, offset) of the label region, or None if there is none (no rows)."""
        return (self.fd, self.offsets['indicator']) if self.shape[0] * self.shape[1] else None

    @property
    def indicator(self):
        """The label region of the file as an array: a plain ndarray over a memory map (its base keeps the map
        alive), so it pickles, copies and compares like any array."""
        if self._map is None:
            if self.shape[0] * self.shape[1]:
                self._map = np.memmap(self.path, dtype=np.uint8, mode='r+', offset=self.offsets['indicator'],
                                      shape=self.shape).view(np.ndarray)
            else:
                self._map = np.empty(self.shape, dtype=np.uint8)
        return self._map

    def _close(self):
        if self.fd is not None:
            os.close(self.fd)
            self.fd = None

    def complete(self, mcweights, mcrates):
        gb = self.gb
        for name, value in (('mcweights', mcweights), ('mcrates', mcrates)):
            arr = np.ascontiguousarray(value, dtype=np.float64)
            if arr.nbytes:
                os.pwrite(self.fd, arr.data, self.offsets[name])
        self._close()
        gb.mcweights, gb.mcrates, gb.indicator = mcweights, mcrates, self.indicator

    def abandon(self):
        """The chain failed: no new file, and the previous one back in place."""
        self._close()
        self._map = None
        try:
            os.remove(self.path)
        except OSError:
            pass
        if self.rotated and os.path.exists(self.path + '.bak'):
            os.rename(self.path + '.bak', self.path)


def dump_reference_pickle(obj, f):
    _RefPickler(f).dump(obj)


def load_reference_pickle(f):
    return _RefUnpickler(f).load()


def _status_error(residue, status):
    if status & _cabi.STATUS_TIMEOUT:
        return _cabi.BrtaError(f'residue {residue}: team rendezvous timed out on the device (status {status})')
    if status != _cabi.STATUS_OK:
        return FloatingPointError(f'residue {residue}: sampler saw a non-finite likelihood (status {status})')
    return None


def _device_count():
    import torch
    if not torch.cuda.is_available():
        raise _cabi.BrtaError('basicrta_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    return torch.cuda.device_count()


def _pick_device(device, loc=0):
    return int(device) if device is not None else int(loc) % _device_count()


WRITER_THREADS = 8
PROGRESS_ROWS = 32             # the kernel publishes finished rows in blocks of this many
LIVE_MIN_BYTES = 64 << 20      # smaller label volumes are not worth a poller thread


def run_batch(gibbs_list, device=None, seed=None, prepared=False, engine=None, save=False,
              keep_indicator_on_device=False, pool=None, progress=None, live=True):
    """Run many residues' chains in ONE launch on one GPU and fill their output arrays.

    All members must share ``ncomp``, ``niter`` and ``g`` (``ParallelGibbs`` guarantees it,
    gibbs.py:73-75).  Hyper-parameters and ``g`` are read here, i.e. at run time.

    Residues fail one by one, as in the reference's pool (one worker's exception does not stop the
    others, gibbs.py:80-88): a residue whose input is unusable is left out of the launch, a chain that
    ends with a non-zero device status is not assigned; all others are assigned (and, with ``save``,
    pickled by a pool of writer threads while the rest of the labels are still coming back over PCIe).
    Failures are reported together at the end as :class:`GibbsBatchError`.

    ``live`` (default): large batches come home WHILE the sweep runs -- the kernel publishes finished row
    blocks, a poller copies them out (``engine.LiveStream``) and, with ``save``, straight into the label region
    of each residue's pickle (:class:`DeferredPickle`; ``indicator`` is then a memory map of the file).
    ``progress(done_iterations, niter)`` is called as the slowest chain of the batch advances.
    """
    if not gibbs_list:
        return
    first = gibbs_list[0]
    key = (first.ncomp, first.niter, first.g)
    failures = []
    members, inputs = [], []
    for gb in gibbs_list:
        if (gb.ncomp, gb.niter, gb.g) != key:
            raise ValueError('a batch needs one (ncomp, niter, g)')
        try:
            if not prepared:
                gb._prepare(allocate_indicator=False)
            inputs.append(gb._chain_input())
            members.append(gb)
        except (ValueError, TypeError, FloatingPointError, MemoryError) as e:
            failures.append((gb.residue, e))
    if members:
        eng = engine if engine is not None else get_engine(_pick_device(device, first.loc))
        if seed is None:
            seed = getattr(first, 'seed', None)
        if seed is None:
            seed = _fresh_seed()
        rows = (first.niter + 1) // first.g
        label_bytes = rows * sum(len(gb.times) for gb in members)
        want_live = bool(live) and not keep_indicator_on_device and rows >= 4 * PROGRESS_ROWS and label_bytes >= LIVE_MIN_BYTES
        nvtx = eng.torch.cuda.nvtx                             # ranges for nsys / ncu timelines
        nvtx.range_push(f'brta prepare {len(members)} residues')
        try:
            db = eng.prepare(inputs, first.ncomp, first.niter, thin=first.g, seed=seed,
                             progress_rows=PROGRESS_ROWS if want_live else 0)
        finally:
            nvtx.range_pop()
        if progress is not None:
            progress(0, first.niter)
        nvtx.range_push('brta sweep')
        eng.launch(db)
        nvtx.range_pop()
        lock = threading.Lock()
        if db.progress is not None:
            _finish_live(eng, db, members, save, pool, progress, failures, lock)
            if progress is not None:
                progress(first.niter, first.niter)
            if failures:
                raise GibbsBatchError(failures)
            return

        def on_chain(r, res):
            gb = members[r]
            err = _status_error(gb.residue, res.status)
            if err is None:
                try:
                    gb.mcweights, gb.mcrates = res.mcweights, res.mcrates
                    if keep_indicator_on_device:
                        gb.indicator, gb._device_indicator, gb._device_batch = None, res.indicator, db
                    else:
                        gb.indicator = res.indicator
                        gb.__dict__.pop('_device_indicator', None)
                    if save:
                        os.makedirs(gb._savedir(), exist_ok=True)
                        gb.save()
                except Exception as e:                      # e.g. disk full: this residue only
                    err = e
            if err is not None:
                with lock:
                    failures.append((gb.residue, err))

        own_pool = pool is None
        pool = ThreadPoolExecutor(max_workers=WRITER_THREADS) if own_pool else pool
        try:
            for fut in eng.stream_results(db, on_chain, pool, keep_on_device=keep_indicator_on_device):
                fut.result()
        finally:
            if own_pool:
                pool.shutdown(wait=True)
        if progress is not None:
            progress(first.niter, first.niter)
    if failures:
        raise GibbsBatchError(failures)


def _finish_live(eng, db, members, save, pool, progress, failures, lock):
    """The overlapped output path of :func:`run_batch`: labels stream into their final place while the launch
    runs; after it, only the last row block, the (tiny) weights / rates and the bookkeeping are left."""
    deferred = [None] * len(members)
    sinks = None
    if save:
        sinks = [None] * len(members)
        for r, gb in enumerate(members):
            try:
                os.makedirs(gb._savedir(), exist_ok=True)
                deferred[r] = DeferredPickle(gb, db.rows)
                sinks[r] = deferred[r].sink
            except Exception as e:                          # e.g. disk full: this residue only (labels go to memory)
                failures.append((gb.residue, e))

    def on_chain(r, res):
        gb = members[r]
        err = _status_error(gb.residue, res.status)
        try:
            if err is None and (not save or deferred[r] is not None):
                if save:
                    deferred[r].complete(res.mcweights, res.mcrates)
                else:
                    gb.mcweights, gb.mcrates, gb.indicator = res.mcweights, res.mcrates, res.indicator
                gb.__dict__.pop('_device_indicator', None)
            elif deferred[r] is not None:
                deferred[r].abandon()
        except Exception as e:
            err = e
        if err is not None:
            with lock:
                failures.append((gb.residue, err))

    own_pool = pool is None
    pool = ThreadPoolExecutor(max_workers=WRITER_THREADS) if own_pool else pool
    try:
        stream = eng.start_live_stream(db, on_chain, pool, progress=progress, sinks=sinks)
        for fut in stream.finish():
            fut.result()
    finally:
        if own_pool:
            pool.shutdown(wait=True)


class ParallelGibbs(object):
    """Run a Gibbs sampler for every residue of a contact map (gibbs.py:20-88).

    ``nproc`` is the number of GPUs to shard the residues over (the reference's number of
    pool processes); whole residues go to GPUs by longest-processing-time-first on N_r --
    chains are independent, so there is no collective.
    """

    def __init__(self, contacts, nproc=1, ncomp=15, niter=110000):
        self.cutoff = float(contacts.strip('.pkl').split('/')[-1].split('_')[-1])
        self.niter = niter
        self.nproc = nproc
        self.ncomp = ncomp
        self.contacts = contacts

    @staticmethod
    def _residue_names(contacts, resids):
        """``W313``-style names from the AtomGroup in the dtype metadata (gibbs.py:62-67);
        plain ``X{resid}`` when the metadata (or MDAnalysis) is unavailable."""
        meta = getattr(contacts.dtype, 'metadata', None) or {}
        try:
            import MDAnalysis as mda
            rg = meta['ag1'].residues
            letters = {int(i): mda.lib.util.convert_aa_code(n) for i, n in zip(rg.resids, rg.resnames)}
            return [f'{letters[int(r)]}{int(r)}' for r in resids]
        except Exception:
            names = meta.get('residue_names') if isinstance(meta, dict) else None
            if names is not None:
                return [str(names[int(r)]) for r in resids]
            return [f'X{int(r)}' for r in resids]

    def _load(self, run_resids):
        return load_contacts(self.contacts, run_resids, self._residue_names)

    def run(self, run_resids=None, skip_existing=False):
        """``skip_existing``: the reference's rerun rule (scripts/get_rerun_residues.py:22-28) -- a residue
        whose ``gibbs_{niter}.pkl`` already exists is not sampled again."""
        resids, names, times = self._load(run_resids)
        n_gpu = max(1, min(int(self.nproc), _device_count()))
        gibbs = [Gibbs(t, name, i % n_gpu, ncomp=self.ncomp, niter=self.niter, cutoff=self.cutoff)
                 for i, (name, t) in enumerate(zip(names, times)) if len(t) > 0]
        if skip_existing:
            gibbs = [g for g in gibbs if not os.path.exists(f'{g._savedir()}/gibbs_{g.niter}.pkl')]
        dispatch(gibbs, n_gpu)
        return gibbs


def load_contacts(path, run_resids=None, namer=None):
    """Per-residue residence times of one ``contacts_{cutoff}.pkl`` (column 0 = protein resid, column 3 =
    duration in ns, contacts.py:227-229): one stable sort groups all residues' durations instead of one
    boolean mask per residue (gibbs.py:68-69 is O(R*M)).  Returns (resids, names, list of times)."""
    with open(path, 'r+b') as f:
        contacts = pickle.load(f)
    protids = np.unique(contacts[:, 0])
    if run_resids is None or (np.ndim(run_resids) == 0 and not run_resids) or \
            (np.ndim(run_resids) > 0 and len(run_resids) == 0):
        run_resids = protids
    if not isinstance(run_resids, (list, np.ndarray)):
        run_resids = [run_resids]
    run_resids = np.asarray(run_resids)
    order = np.argsort(contacts[:, 0], kind='stable')
    col0 = np.asarray(contacts[:, 0])[order]
    dur = np.asarray(contacts[:, 3], dtype=np.float64)[order]
    lo = np.searchsorted(col0, run_resids, side='left')
    hi = np.searchsorted(col0, run_resids, side='right')
    times = [dur[a:b].copy() for a, b in zip(lo, hi)]
    namer = namer or ParallelGibbs._residue_names
    return run_resids, namer(contacts, run_resids), times


def dispatch(gibbs_list, n_gpu, seed=None, save=True, keep_indicator_on_device=False, progress=None):
    """Shard residues over GPUs (one host thread per GPU, one launch per GPU) and write the
    reference's per-residue pickles.  ``gibbs_list`` may mix batches -- residues of several contact
    cutoffs (five ``contacts_{cutoff}.pkl`` -> one launch per GPU), as long as
    they share (ncomp, niter, g).  Failed residues are collected over all GPUs and raised together as
    :class:`GibbsBatchError` after every other residue has been written."""
    if not gibbs_list:
        return
    shards = shard_chains([len(g.times) for g in gibbs_list], n_gpu)
    failures, errors = [], []
    seed = _fresh_seed() if seed is None else seed
    pool = ThreadPoolExecutor(max_workers=WRITER_THREADS * max(1, min(n_gpu, 2)))

    def work(dev, idx):
        try:
            members = [gibbs_list[i] for i in idx]
            for gb in members:
                gb.loc = dev
            run_batch(members, device=dev, seed=seed, save=save, pool=pool,
                      keep_indicator_on_device=keep_indicator_on_device, progress=progress)
        except GibbsBatchError as e:
            failures.extend(e.failures)
        except BaseException as e:                     # surfaced to the caller below
            errors.append(e)

    threads = [threading.Thread(target=work, args=(dev, idx)) for dev, idx in enumerate(shards) if len(idx)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    pool.shutdown(wait=True)
    if errors:
        raise errors[0]
    if failures:
        raise GibbsBatchError(failures)


class MultiCutoffGibbs(object):
    """The multi-cutoff ensemble (BASELINE.json config 3): the same protein at several contact cutoffs,
    i.e. several ``contacts_{cutoff}.pkl`` files of ``contacts.ProcessContacts`` (contacts.py:181-191).
    The reference runs ``python -m basicrta.gibbs --contacts ...`` once per file; here the residues of
    ALL files go to the GPUs as one batch per GPU (2 000 chains for 400 residues x 5 cutoffs), and every
    residue is written to its own ``basicrta-{cutoff}/{residue}/`` as before."""

    def __init__(self, contact_files, nproc=1, ncomp=15, niter=110000):
        self.members = [ParallelGibbs(c, nproc=nproc, ncomp=ncomp, niter=niter) for c in contact_files]
        self.nproc, self.ncomp, self.niter = nproc, ncomp, niter

    def run(self, run_resids=None, skip_existing=False):
        n_gpu = max(1, min(int(self.nproc), _device_count()))
        gibbs = []
        for pg in self.members:
            resids, names, times = pg._load(run_resids)
            gibbs += [Gibbs(t, name, 0, ncomp=self.ncomp, niter=self.niter, cutoff=pg.cutoff)
                      for name, t in zip(names, times) if len(t) > 0]
        if skip_existing:
            gibbs = [g for g in gibbs if not os.path.exists(f'{g._savedir()}/gibbs_{g.niter}.pkl')]
        dispatch(gibbs, n_gpu)
        return gibbs


if __name__ == '__main__':                             # gibbs.py:781-795
    import argparse
    parser = argparse.ArgumentParser()
    parser.add_argument('--contacts', nargs='+')
    parser.add_argument('--resid', type=int, default=None)
    parser.add_argument('--nproc', type=int, default=1)
    parser.add_argument('--niter', type=int, default=110000)
    parser.add_argument('--ncomp', type=int, default=15)
    parser.add_argument('--skip-existing', action='store_true')
    args = parser.parse_args()
    if len(args.contacts) == 1:
        ParallelGibbs(args.contacts[0], nproc=args.nproc, ncomp=args.ncomp, niter=args.niter).run(
            run_resids=args.resid, skip_existing=args.skip_existing)
    else:
        MultiCutoffGibbs(args.contacts, nproc=args.nproc, ncomp=args.ncomp, niter=args.niter).run(
            run_resids=args.resid, skip_existing=args.skip_existing)
