#!/usr/bin/env python
"""Benchmark of the likelihood hot path (BASELINE.json configs[1]):
batched photometric log-likelihood, 2^20 live points x 10 bands per GPU, on a
synthetic MIST grid of the real shape (607 x 140 x 19 x 5 x 9) and a random-init
photometric ANN of the assumed real architecture (10 x (6 -> 128 -> 128 -> 1)).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path (MIST interpolation -> parameter assembly ->
photometric MLP + chi-square -> lnL) over one batch of theta already resident in
HBM.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement".
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))

METRIC = 'phot_loglike_evals_per_sec'
UNIT = 'evals/s'
NAMES = ['Teff', 'log(g)', '[Fe/H]', '[a/Fe]', 'Vrad', 'Vrot', 'Vmic', 'Inst_R', 'log(R)', 'Dist', 'Av',
         'log(A)', 'EEP', 'initial_[Fe/H]', 'initial_[a/Fe]', 'initial_Mass']
PHOT_ON = {'Dist', 'Av', 'EEP', 'initial_[Fe/H]', 'initial_[a/Fe]', 'initial_Mass'}
KIND_TEXT = {'reference': "the reference's own likelihood.likelihood class, unmodified modules from oracle/_ref",
             'port': 'the oracle port oracle/like_port.py (no reference copy present)'}
PARITY_ROWS = 4096
SPEC_PARITY_ROWS = 48


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default=os.environ.get('MS_BENCH_PRECISION', 'tc'))
    ap.add_argument('--batch', type=int, default=1 << 20, help='live points per GPU')
    ap.add_argument('--hidden', type=int, default=128)
    ap.add_argument('--small-grid', action='store_true', help='debug: 60x24x7x5 grid')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--spec-batch', type=int, default=65536,
                    help='points per GPU of the auxiliary spectrum + photometry measurement (0 = skip)')
    ap.add_argument('--catalog-stars', type=int, default=100000,
                    help='stars per GPU fitted with the batched nested sampler for the stars/hour figure (0 = skip); '
                         'the default is the size of configs[4]')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sust=d['bf16_tflops_sustained'],
                    src='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src='fallback')


def make_workload(args):
    from minesweeper_b200 import synth
    if args.small_grid:
        mist = synth.make_mist(ne=60, nm=24, nf=7, na=5, seed=0)
        box = dict(Dist=(1.0, 100.0), Av=(0.0, 0.1), EEP=(202.0, 261.0), feh=(-4.0, 0.0),
                   afe=(-0.2, 0.6), mass=(0.5, 1.25))
    else:
        mist = synth.make_mist(seed=0)
        box = None
    phot = synth.make_phot_net(h=args.hidden, seed=1)
    fitpars = [NAMES, {n: (n in PHOT_ON) for n in NAMES}]
    fitargs = dict(ageweight=True, mistdif=True, fixedpars={}, MISTpath=mist, photANNpath=phot,
                   obs_phot=synth.demo_obs_phot())
    runbools = [False, True, False, False, False]
    return fitargs, fitpars, runbools, box


def workload_name(args, mist):
    g = 'x'.join(str(len(mist[k])) for k in ('eep', 'mass', 'feh', 'afe'))
    return ('configs[1]: batched photometric loglike, %d live points x 10 bands per GPU, MIST grid %s x 9, '
            'phot ANN 10 x (6-%d-%d-1) sigmoid' % (args.batch, g, args.hidden, args.hidden))


class ClockSampler(object):
    """Samples SM clock / power / throttle reasons of one GPU through NVML from a
    thread (every ~2 ms) while the timed region runs."""

    def __init__(self, index):
        self.rows, self.ok, self.stop_flag = [], False, False
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            self.ok = True
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception as e:                     # pragma: no cover
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0,
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                                  if hasattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons')
                                  else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        if not self.ok:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvml unavailable'])
        self.stop_flag = True
        self.t.join(1.0)
        nv = self.nv
        names = {'hw_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8),
                 'hw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40),
                 'sw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20),
                 'sw_power_cap': getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4)}
        sm = [r[0] for r in self.rows]
        reasons = sorted(k for k, bit in names.items() if any(r[2] & bit for r in self.rows))
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=float(self.max_mhz),
                    power_w_max=max((r[1] for r in self.rows), default=None), samples=len(sm),
                    reasons=reasons)


def parity_check(L, eng, theta_np, lnl_timed, ref, precision):
    """The timed step's own lnL (one-kernel tensor-core form in 'tc') on the first rows of the batch against the
    oracle: -inf rows identical, lnL within the per-row chi-square bound of a 1e-5 relative magnitude error (the
    north_star's fp32 tolerance; 1e-6 absolute in f64); brackets bit-exact (two-kernel form exposes them) and
    magnitudes 1e-5 relative at the oracle's interpolated parameters.  Raises on failure."""
    n = len(ref['lnl'])
    ok = ref['ok']
    got = lnl_timed[:n]
    assert np.array_equal(np.isneginf(got), ~ok), 'parity: out-of-grid rows differ from the oracle'
    err = np.abs(got[ok] - ref['lnl'][ok])
    mag_tol = {'f64': 1e-11, 'tc16': 3e-3}.get(precision, 1e-5)
    if precision == 'f64':
        bound = np.full(ok.sum(), 1e-6)
    elif precision == 'tc16':            # stated absolute magnitude tolerance (3e-3 mag) instead of 1e-5 relative
        bound = np.nansum((np.abs(ref['mags'][ok] - ref['obs_mag']) * 3e-3 + 0.5 * 3e-3 ** 2) / ref['obs_err'] ** 2,
                          axis=1) + 1e-6
    else:
        bound = ref['bound'][ok] + 1e-6 + 1e-9 * np.abs(ref['lnl'][ok])
    eng.set_fusion(False)
    _, _, br = L.lnlike_batch(theta_np[:n], want_brackets=True)
    eng.set_fusion(True)
    br_ok = bool(np.array_equal(br.cpu().numpy(), ref['brackets']))
    mags = eng.phot_sed(ref['photpars'][ok]).cpu().numpy()
    rel_mag = np.abs(mags - ref['mags'][ok]) / np.maximum(np.abs(ref['mags'][ok]), 1.0)
    out = dict(rows=int(n), out_of_grid_rows=int((~ok).sum()), brackets_bit_exact=br_ok,
               max_rel_mag=float(rel_mag.max()), max_dlnL=float(err.max()),
               max_dlnL_over_bound=float((err / bound).max()), mag_tolerance=mag_tol,
               checked='lnL of the timed kernel vs oracle.like_port.lnlike_phot_dense on the first rows of the batch')
    assert br_ok, 'parity: brackets differ from the oracle'
    assert rel_mag.max() <= mag_tol, 'parity: magnitudes off by %g' % rel_mag.max()
    assert np.all(err <= bound), 'parity: lnL outside the per-row bound (%g x)' % (err / bound).max()
    return out


def run_reference(args):
    """CPU arm: the reference's per-point execution model on all host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from minesweeper_b200 import synth
    from oracle.cpu_baseline import CpuArm
    fitargs, fitpars, runbools, box = make_workload(args)
    theta = synth.draw_theta_phot(1 << 16, seed=3, box=box)
    arm = CpuArm(fitargs, fitpars, runbools, theta)
    r1 = arm.calibrate(0.5)
    budget = 120.0
    per_step = int(max(r1 * arm.cores * 0.8 * budget / (args.steps + args.warmup), 64 * arm.cores))
    for _ in range(args.warmup):
        arm.step(per_step)
    n_tot, t_tot = 0, 0.0
    for _ in range(args.steps):
        n, t = arm.step(per_step)
        n_tot += n; t_tot += t
    arm.close()
    val = n_tot / t_tot
    sample = ('%d evaluations per step over %d worker processes (per-point lnlikefn loop of %s)'
              % (per_step, arm.cores, KIND_TEXT[arm.kind]))
    print(json.dumps(dict(
        impl='reference', metric=METRIC, value=val, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
        warmup=args.warmup, ms_per_step=1e3 * t_tot / args.steps, higher_is_better=True, scaling='weak',
        vs_baseline=None, dtype='f64', data='synthetic',
        config=dict(workload=workload_name(args, fitargs['MISTpath']), sample=sample),
        cpu_baseline=dict(value=val, unit=UNIT, cores=arm.cores, kind=arm.kind, sample=sample),
        e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
        gpu_launches=0)))


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    fitargs, fitpars, runbools, box = make_workload(args)
    from minesweeper_b200 import synth

    # ---- CPU baseline first: its workers are forked, so CUDA must not be initialised yet
    cpu = None
    cpu_rate1 = cpu_kind = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.cpu_baseline import CpuArm
        arm = CpuArm(fitargs, fitpars, runbools, synth.draw_theta_phot(1 << 16, seed=3, box=box))
        r1 = arm.calibrate(0.5)
        n = int(max(r1 * arm.cores * 0.8 * args.cpu_seconds, 64 * arm.cores))
        arm.step(max(n // 20, arm.cores))
        ne, te = arm.step(n)
        arm.close()
        cpu = dict(value=ne / te, unit=UNIT, cores=arm.cores, kind=arm.kind,
                   sample='%d per-point lnlikefn evaluations of the same workload over %d worker '
                          'processes (%.1f s; %s); single-core probe %.0f evals/s'
                          % (ne, arm.cores, te, KIND_TEXT[arm.kind], r1))
        cpu_rate1, cpu_kind = r1, arm.kind
        del arm

    # ---- oracle values for the in-run parity sample (CPU leg): the first PARITY_ROWS rows of THIS rank-0 batch,
    # 1 % of them out of grid, through the dense oracle (pinned to the reference by tests/golden)
    parity_ref = None
    if rank == 0:
        from oracle.like_port import lnlike_phot_dense
        parity_ref = lnlike_phot_dense(fitargs['MISTpath'], fitargs['photANNpath'], fitargs['obs_phot'],
                                       synth.draw_theta_phot(args.batch, seed=3, box=box)[:PARITY_ROWS])

    # ---- CPU leg of the spectrum + photometry configuration (configs[2]): the reference's per-point loop on all
    # cores, and its lnL on the first rows of the batch as the parity reference of the GPU run below
    spec_cpu = spec_parity_ref = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.spec_batch > 0:
        from bench_spec import make_spec_workload
        from oracle.cpu_baseline import CpuArm
        sfa, sfp, srb, sth = make_spec_workload(fitargs['MISTpath'], fitargs['photANNpath'], args.spec_batch, 1536,
                                                seed=rank)
        arm = CpuArm(sfa, sfp, srb, sth[:8192])
        r1 = arm.calibrate(1.0)
        # the spectrum path is memory-bound on the host (24 MB of output-layer weights per evaluation): the workers do
        # not scale like the single-core probe, so the sample is sized from a short parallel step
        n0, t0_ = arm.step(4 * arm.cores)
        n = int(max(n0 / t0_ * args.cpu_seconds * 0.6, 4 * arm.cores))
        ne, te = arm.step(n)
        rows = np.arange(SPEC_PARITY_ROWS)
        lref = arm.eval_rows(rows)
        arm.close()
        spec_cpu = dict(value=ne / te, unit=UNIT, cores=arm.cores, kind=arm.kind,
                        sample='%d per-point lnlikefn evaluations (spectrum 1536 px + 10 bands) over %d worker '
                               'processes (%.1f s; %s); single-core probe %.0f evals/s'
                               % (ne, arm.cores, te, KIND_TEXT[arm.kind], r1))
        # |d lnL| a 1e-5 relative model-flux error can cause at S/N 150 (chi-square sensitivity), + the photometric part
        tol = 1e-5 * 150.0 * np.sqrt(2.0 * 2.0 * np.abs(lref) * 1536) + 1e-5 * np.abs(lref) + 1e-3
        spec_parity_ref = (len(rows), lref, tol)
        del arm

    # NCCL prints its version banner on stdout at NCCL_DEBUG=VERSION (read when the library loads); stdout carries
    # the one JSON line
    if world > 1 and os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
        os.environ['NCCL_DEBUG'] = 'WARN'
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa_cores = None
    if world > 1:
        from minesweeper_b200.distributed import bind_to_gpu_numa
        numa_cores = bind_to_gpu_numa(local)          # before the pinned staging buffers are allocated
        dist.init_process_group('nccl', device_id=dev)
    from minesweeper_b200.likelihood import likelihood
    L = likelihood(fitargs, fitpars, runbools, precision=args.precision, device=local, verbose=False)
    eng = L.engine
    B = args.batch
    theta_np = synth.draw_theta_phot(B, seed=3 + rank, box=box)
    theta = torch.from_numpy(theta_np).to(dev)
    # Results are double-buffered.  The one collective of the path (the gather of lnL) is fused into the likelihood
    # kernel: its epilogue stores lnL into every rank's gather buffer through peer-mapped symmetric memory
    # (distributed.PeerGather); only a small symmetric-memory barrier follows, on a side stream, overlapping the next
    # step.  Fallback when symmetric memory is unavailable: all_gather_into_tensor (NCCL) on the side stream.
    lnl_bufs = [torch.empty(B, dtype=torch.float64, device=dev) for _ in range(2)]
    peer = None
    gathered = None
    gather_kind = 'none'
    if world > 1:
        try:
            if os.environ.get('MS_BENCH_GATHER', 'peer') != 'peer' or args.precision not in ('tc', 'tc16'):
                raise RuntimeError('peer gather not selected')
            from minesweeper_b200.distributed import PeerGather
            peer = PeerGather(B, dev)
            gather_kind = ('fused into the photometry kernel: its epilogue stores lnL into every rank\'s gather buffer over '
                           'NVLink (peer-mapped symmetric memory, %d bytes per rank per step out and in); a '
                           'symmetric-memory barrier on a side stream orders the ranks' % peer.bytes_per_rank_per_step)
        except Exception as e:
            peer = None
            gathered = [torch.empty(B * world, dtype=torch.float64, device=dev) for _ in range(2)]
            gather_kind = 'all_gather_into_tensor(lnL) per step (NCCL), on a side stream, overlapping the next step [%s]' % (
                type(e).__name__)
    comm = torch.cuda.Stream(dev) if world > 1 else None
    gather_done = [None, None]
    nstep = [0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def step():
        i = nstep[0] & 1
        nstep[0] += 1
        main = torch.cuda.current_stream()
        if gather_done[i] is not None:
            main.wait_event(gather_done[i])                             # this buffer's previous gather is complete
        if peer is not None:
            peer.arm(eng, i)
        eng.lnlike_batch(theta, L._colmap, L._fixed, L._flags, want_derived=False, out=lnl_bufs[i])
        if world > 1:
            ready = torch.cuda.Event()
            ready.record(main)
            with torch.cuda.stream(comm):
                comm.wait_event(ready)
                if peer is not None:
                    peer.complete(i)                                    # every rank's stores into buffer i have landed
                else:
                    dist.all_gather_into_tensor(gathered[i], lnl_bufs[i])
                gather_done[i] = torch.cuda.Event()
                gather_done[i].record(comm)

    def drain():
        for e in gather_done:
            if e is not None:
                torch.cuda.current_stream().wait_event(e)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    eng.profile(True)
    l0 = eng.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for k, (e0, e1) in enumerate(ev):
        flush.zero_()                                                   # L2 flush, outside the timed pair
        e0.record()
        step()
        if k == len(ev) - 1:
            drain()                                                     # every gather is inside the timed region
        e1.record()
    barrier()
    lnl = lnl_bufs[(nstep[0] - 1) & 1]
    if peer is not None:
        # the fused gather delivered this rank's block to itself and every peer's block here
        gi = (nstep[0] - 1) & 1
        torch.cuda.synchronize()
        dist.barrier()
        gat = peer.gathered(gi)
        assert torch.equal(gat[rank * B:(rank + 1) * B], lnl), 'fused gather: own block differs'
        chk = torch.stack([gat[r * B:(r + 1) * B][torch.isfinite(gat[r * B:(r + 1) * B])].sum() for r in range(world)])
        ref = torch.zeros(world, dtype=torch.float64, device=dev)
        ref[rank] = lnl[torch.isfinite(lnl)].sum()
        dist.all_reduce(ref)
        assert torch.allclose(chk, ref, rtol=1e-12), 'fused gather: a peer block differs from its owner\'s result'
        eng.set_lnl_peers()
    ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    launches = eng.launch_count() - l0
    prof = eng.profile_read()
    eng.profile(False)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = B * world * args.steps / (ms * 1e-3)

    # ---- second timed variant: lnL AND the derived[B, 13] block the prior consumes
    der_buf = torch.empty((B, 13), dtype=torch.float64, device=dev)
    for _ in range(3):
        eng.lnlike_batch(theta, L._colmap, L._fixed, L._flags, want_derived=True, out=lnl_bufs[0], derived_out=der_buf)
    barrier()
    dv = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for e0, e1 in dv:
        flush.zero_()
        e0.record()
        eng.lnlike_batch(theta, L._colmap, L._fixed, L._flags, want_derived=True, out=lnl_bufs[0], derived_out=der_buf)
        e1.record()
    barrier()
    tdv = torch.tensor([sum(e0.elapsed_time(e1) for e0, e1 in dv)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tdv, op=dist.ReduceOp.MAX)
    derived_variant = dict(value=B * world * args.steps / (float(tdv.item()) * 1e-3), unit=UNIT,
                           ms_per_step=float(tdv.item()) / args.steps,
                           extra_bytes_per_eval=13 * 8, note='same step, also writing derived[B,13] f64 to HBM')
    del der_buf

    # ---- end to end through the host-buffer C-ABI call: pinned host theta in, lnL out, every step
    th_pin = torch.from_numpy(theta_np).pin_memory()
    out_pin = torch.empty(B, dtype=torch.float64).pin_memory()
    th_h, out_h = th_pin.numpy(), out_pin.numpy()
    for _ in range(3):
        L.lnlike_batch_host(th_h, lnl_out=out_h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        L.lnlike_batch_host(th_h, lnl_out=out_h)
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = B * world * args.steps / float(te.item())
    # the host entry point evaluates chunk by chunk; a short last chunk takes the two-kernel form (f64 corner sums), the
    # one-kernel form sums the corners in f32: they agree to the fp32-class magnitude tolerance (1e-5 mag in every band)
    dev_l = lnl.cpu().numpy()
    fin_ = np.isfinite(dev_l)
    sig_min = min(float(v[1]) for v in fitargs['obs_phot'].values())
    tol_hd = 1e-5 * np.sqrt(2.0 * np.abs(dev_l[fin_]) * len(fitargs['obs_phot'])) / sig_min + 1e-6
    assert np.array_equal(np.isfinite(out_h), fin_) and np.all(np.abs(out_h[fin_] - dev_l[fin_]) <= tol_hd), \
        'host and device entry points disagree'

    # ---- in-run parity: the output of the timed (fused) kernel itself against the oracle sample
    parity = None
    if parity_ref is not None:
        parity = parity_check(L, eng, theta_np, dev_l, parity_ref, args.precision)

    # ---- stars fitted per hour (BASELINE.json metric, configs[4]).  Headline = the shortest random walk that passes
    # the accuracy gate against brute-force quadrature (tools/bench_catalog.py: GATE, GATE_WALKS); the demo's own
    # setting (walks = 5, demo/etc/samplerinfo.dat) is timed on a 20 % subsample and reported with its gate result.
    catalog = None
    if args.catalog_stars > 0:
        try:
            from bench_catalog import make_catalog_sampler, catalog_priordict, accuracy_block, GATE_WALKS
            from minesweeper_b200.prior import prior
            cpd = catalog_priordict(261.0 if args.small_grid else 808.0)
            Pc = prior(fitargs, cpd, fitpars, runbools, likeobj=L)

            def run_catalog(nstars, walks):
                ns, tht = make_catalog_sampler(L, Pc, nstars, dev, 150, walks, 0.01, 20000, rank)
                barrier()
                tc0 = time.perf_counter()
                rc = ns.run(check_every=16, compact_frac=0.85)      # measured: 15.1 s (64, 0.5) -> 13.5 s per 100 k stars
                torch.cuda.synchronize()
                tcat = torch.tensor([time.perf_counter() - tc0], dtype=torch.float64, device=dev)
                ncalls = torch.tensor([float(rc['ncall'].sum())], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(tcat, op=dist.ReduceOp.MAX)
                    dist.all_reduce(ncalls)
                acc = accuracy_block(L, Pc, ns, rc, tht, cpd, 64) if rank == 0 else None
                secs = float(tcat.item())
                return dict(stars=nstars * world, seconds=secs, stars_per_hour=nstars * world / secs * 3600.0,
                            loglike_calls=float(ncalls.item()), loglike_calls_per_sec=float(ncalls.item()) / secs,
                            nlive=150, walks=walks, dlogz=0.01, queue=ns.Q, converged_frac=float(rc['converged'].mean()),
                            accuracy=acc)

            catalog = run_catalog(args.catalog_stars, GATE_WALKS)
            catalog['sampler'] = ('lock-step static nested sampling, rwalk, %d queued walkers per star (sampler.auto_queue; up to 64 for the '
                                  'stragglers), bookkeeping in CUDA ' % catalog['queue'] +
                                  'kernels, CUDA-graph replay')
            catalog['note'] = 'configs[4]: stars per GPU fixed (weak scaling); walks = shortest walk passing accuracy.gate'
            catalog['demo_settings'] = run_catalog(max(args.catalog_stars // 5, 64), 5)
            # configs[3]: ONE star fitted on its own (64 queued walkers), a few stars one after the other; next to it
            # what the same fit costs the reference's one-point-at-a-time loop on one host core (dynesty itself is not
            # installed: iterations x walks + nlive sequential lnlikefn calls at the measured single-core rate)
            if world == 1:
                fits = []
                for i in range(6):
                    ns1, tht1 = make_catalog_sampler(L, Pc, 1, dev, 150, GATE_WALKS, 0.01, 20000, 1000 + i)
                    torch.cuda.synchronize()
                    tf0 = time.perf_counter()
                    r1f = ns1.run()
                    torch.cuda.synchronize()
                    tfit = time.perf_counter() - tf0
                    a1 = accuracy_block(L, Pc, ns1, r1f, tht1, cpd, 1)
                    fits.append(dict(seconds=tfit, iterations=int(r1f['niter'][0]), rounds=int(r1f['rounds']),
                                     loglike_calls=int(r1f['ncall'][0]), logz=float(r1f['logz'][0]),
                                     logzerr=float(r1f['logzerr'][0]), dlnz_vs_bruteforce=a1['dlnz_median'],
                                     max_mean_bias_in_sigma=max(a1['mean_bias_median_abs_in_sigma'].values()),
                                     converged=bool(r1f['converged'][0])))
                med = sorted(f['seconds'] for f in fits)[len(fits) // 2]
                seq = [f['iterations'] * GATE_WALKS + 150 for f in fits]
                catalog['single_fit'] = dict(
                    workload='configs[3]: one full nested-sampling fit of one star (10 bands + parallax prior; static, rwalk, '
                             'nlive 150, walks %d, dlogz 0.01), %d stars fitted one after the other' % (GATE_WALKS, len(fits)),
                    seconds_median=med, queue=ns1.Q, fits=fits,
                    cpu_seconds_estimate_1core=(float(np.median(seq)) / cpu_rate1) if cpu_rate1 else None,
                    cpu_note='median (iterations x walks + nlive) = %d sequential lnlikefn calls at the single-core rate of '
                             'the CPU arm (%s); prior and sampler overhead of dynesty not included'
                             % (int(np.median(seq)), KIND_TEXT.get(cpu_kind, 'not measured')))
        except Exception as e:                                           # never lose the main line
            catalog = dict(error=repr(e))

    # ---- spectrum + photometry likelihood (BASELINE.json metric, configs[2]): 1536 observed pixels (the demo
    # spectrum) with roofline, CPU arm and an in-run parity sample, and the 10 000-pixel variant of the config text
    specblk = None
    if args.spec_batch > 0:
        try:
            from bench_spec import run_spec_bench
            sprec = args.precision if args.precision != 'f64' else 'f32'
            r = run_spec_bench(fitargs['MISTpath'], fitargs['photANNpath'], batch=args.spec_batch, nobs=1536,
                               precision=sprec, steps=3, device=local, seed=rank, parity_ref=spec_parity_ref)
            r10 = run_spec_bench(fitargs['MISTpath'], fitargs['photANNpath'], batch=args.spec_batch, nobs=10000,
                                 precision=sprec, steps=3, device=local, seed=rank)
            tsp = torch.tensor([r['ms_per_step'], r10['ms_per_step']], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tsp, op=dist.ReduceOp.MAX)
            ms15, ms10 = float(tsp[0].item()), float(tsp[1].item())
            val = args.spec_batch * world / (ms15 * 1e-3)
            pk = peaks()
            tf_step = r['flop_per_eval'] * val / world / 1e12            # per GPU, whole step
            specblk = dict(value=val, unit='evals/s', ms_per_step=ms15, batch_per_gpu=args.spec_batch,
                           workload=r['workload'], kernel_ms=r['kernel_ms'],
                           roofline=dict(bound='tensor', kernel='whole spectrum + photometry step (spectral MLP on tcgen05; '
                                         'the broadening kernel, CUDA-core FFT + FIR, is the long pole)',
                                         flop_per_eval=r['flop_per_eval'], achieved=tf_step, peak=pk['tf_burst'],
                                         unit='TFLOP/s', frac=tf_step / pk['tf_burst'], peak_source=pk['src'],
                                         mlp_kernels_alone_tflops=r['spec_mlp_tflops'],
                                         mlp_kernels_alone_frac=r['spec_mlp_tflops'] / pk['tf_burst'],
                                         share_of_step={k: v / ms15 for k, v in r['kernel_ms'].items()}),
                           cpu_baseline=spec_cpu, parity_check=r['parity_check'],
                           n_obs_10000=dict(value=args.spec_batch * world / (ms10 * 1e-3), unit='evals/s',
                                            ms_per_step=ms10, kernel_ms=r10['kernel_ms'], workload=r10['workload']))
        except Exception as e:
            specblk = dict(error=repr(e))

    # ---- other hidden widths of the photometric nets ("default H_p = 128; also report 64 and
    # 256"): same workload, same timing loop, same fused tcgen05 kernel family (256: one tile per CTA, layer-2 weights
    # streamed in K slices)
    hidden_variants = {}
    if args.hidden == 128 and not args.small_grid and args.precision == 'tc':
        for hh, hprec in ((64, 'tc'), (256, 'tc')):
            try:
                fa_h = dict(fitargs, photANNpath=synth.make_phot_net(h=hh, seed=1))
                Lh = likelihood(fa_h, fitpars, runbools, precision=hprec, device=local, verbose=False)
                outh = torch.empty(B, dtype=torch.float64, device=dev)
                nst = 5
                for _ in range(2):
                    Lh.engine.lnlike_batch(theta, Lh._colmap, Lh._fixed, Lh._flags, want_derived=False, out=outh)
                evh = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nst)]
                for e0, e1 in evh:
                    flush.zero_()
                    e0.record()
                    Lh.engine.lnlike_batch(theta, Lh._colmap, Lh._fixed, Lh._flags, want_derived=False, out=outh)
                    e1.record()
                torch.cuda.synchronize()
                msh = sum(e0.elapsed_time(e1) for e0, e1 in evh) / nst
                flop_h = len(Lh.engine.bands) * 2.0 * (6 * hh + hh * hh + hh)
                hidden_variants[str(hh)] = dict(value=B / (msh * 1e-3), unit=UNIT, ms_per_step=msh, precision=hprec,
                                                flop_per_eval=flop_h, achieved_tflops=flop_h * B / (msh * 1e-3) / 1e12,
                                                kernel='phot_mlp_tc (tcgen05, fused)' if hprec == 'tc' else 'phot_mlp_kernel<float> (CUDA cores)')
                Lh.engine.close()
                del Lh, outh
            except Exception as e:
                hidden_variants[str(hh)] = dict(error=repr(e))

    # ---- the two kernels on their own (the timed step runs them fused in the tensor-core modes)
    eng.set_fusion(False)
    for _ in range(2):
        step()
    eng.profile(True)
    for _ in range(5):
        flush.zero_()
        step()
    prof2 = eng.profile_read()
    eng.profile(False)
    eng.set_fusion(True)

    if rank == 0:
        pk = peaks()
        H, nb = args.hidden, len(eng.bands)
        flop_pt = nb * 2.0 * (6 * H + H * H + H)
        ph_ms, ph_n = prof['phot']
        in_ms, in_n = prof2['interp']
        fused = prof['interp'][1] == 0 and args.precision in ('tc', 'tc16')
        achieved = flop_pt * B * ph_n / (ph_ms * 1e-3) / 1e12 if ph_ms > 0 else 0.0
        # DRAM bytes per launch from the last `ncu --set full` capture (tools/traffic_from_ncu.py), with a check that the
        # kernel's sources are still the ones that were captured
        traffic, traffic_capture = None, None
        tp = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
        if os.path.exists(tp):
            tj = json.load(open(tp))
            tkey = ('phot_fused_' if fused else 'phot_') + args.precision
            traffic = tj.get(tkey)
            cap = tj.get(tkey + '_capture')
            if cap:
                import hashlib
                hsh = hashlib.sha256()
                try:
                    for f in cap['sources']:
                        hsh.update(open(os.path.join(ROOT, 'minesweeper_b200', 'csrc', f), 'rb').read())
                    same = hsh.hexdigest() == cap['sources_sha256']
                except OSError:
                    same = None
                traffic_capture = dict(report=cap['report'], kernel_sources_unchanged_since_capture=same)
        tbl_b = 4 if args.precision != 'f64' else 8
        # theta row in, 16 corners x 9 columns gathered, 20-double parameter block out (fused kernel)
        interp_bytes = theta_np.shape[1] * 8 + 16 * 9 * tbl_b + 20 * 8
        out = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
            ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
            dtype={'f64': 'f64', 'f32': 'f32', 'tc': 'f16x3 (split fp16, fp32 accumulate)',
                   'tc16': 'f16 (fp32 accumulate)'}[args.precision], data='synthetic',
            config=dict(workload=workload_name(args, fitargs['MISTpath']), precision=args.precision,
                        l2='flushed between steps (256 MiB memset outside the timed events)',
                        collective=gather_kind,
                        host_affinity=('%d cores local to the GPU' % len(numa_cores)) if numa_cores else 'default',
                        flop_per_eval=flop_pt),
            clocks=clocks,
            e2e=dict(value=e2e, unit=UNIT, h2d_bytes_per_step=B * theta_np.shape[1] * 8,
                     d2h_bytes_per_step=B * 8),
            gpu_launches=int(launches),
            roofline=dict(bound='tensor', kernel='phot_mlp_tc (MIST interpolation fused into its prologue)' if fused
                          else 'phot_mlp', achieved=achieved, peak=pk['tf_burst'],
                          unit='TFLOP/s', frac=achieved / pk['tf_burst'], traffic=traffic,
                          traffic_capture=traffic_capture,
                          peak_source=pk['src'], ms_per_launch=ph_ms / max(ph_n, 1),
                          share_of_step=ph_ms / ms if ms else None),
            roofline_interp=dict(bound='hbm', kernel='mist_interp (on its own: two-kernel sequence, 5 extra steps)',
                                 achieved=interp_bytes * B * in_n / (in_ms * 1e-3) / 1e9 if in_ms > 0 else 0.0,
                                 peak=pk['hbm'], unit='GB/s', bytes_per_eval=interp_bytes,
                                 frac=(interp_bytes * B * in_n / (in_ms * 1e-3) / 1e9 / pk['hbm']) if in_ms > 0 else 0.0,
                                 ms_per_launch=in_ms / max(in_n, 1)),
            kernel_ms={k: v[0] / args.steps for k, v in prof.items() if v[1]},
            kernel_ms_two_kernel_sequence={k: v[0] / 5 for k, v in prof2.items() if v[1]},
            parity_check=parity,
            with_derived=derived_variant,
            hidden_widths=hidden_variants,
            catalog=catalog,
            spec_phot=specblk,
            cpu_baseline=cpu)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
