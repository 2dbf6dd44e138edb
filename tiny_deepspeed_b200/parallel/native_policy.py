"""NativePolicy — the B200 product path of DDP / ZeRO-1 / ZeRO-2 / ZeRO-3.

Parameters and gradients live in flat *symmetric* buffers (same offsets on every rank, peer-mapped,
multicast-bound).  The public ``{name: rank}`` map stays the user-facing artefact; here it is a set of
views over those buffers.

* **DDP** — backward GEMMs write dW straight into the symmetric gradient buffer; as soon as a bucket of
  consecutive tensors is complete a two-shot NVLS all-reduce kernel (``multimem.ld_reduce`` +
  ``multimem.st``) runs on the communication stream while backward continues.  The last bucket (the
  embedding gradients, produced by the final backward kernel) is hidden under the Adam update of everything
  that was reduced earlier (``flush_async`` / ``join``).
* **ZeRO-1/2** — the reference's move (collective right after each gradient, `zero1/module.py:17-24`) at bucket
  granularity: as soon as a bucket of gradients is complete, ONE fused kernel on the communication stream —
  switch-reduced gradient → scale → Adam on the owner's fp32 master/moments → bf16 parameter multicast to all ranks
  (``csrc/comm_sm100.cu: zero_fused_adam_kernel``) — runs underneath the rest of backward.  ZeRO-1 keeps the
  full-size gradient buffer; ZeRO-2/3 write gradients into a small symmetric *ring* of bucket-sized slots that is
  recycled as backward proceeds (slot b is reused by bucket b + nslots once every rank's step kernel for bucket b has
  finished), so a non-owner holds at most `nslots` buckets of gradients (`zero2/module.py:26-36` drops them one by one).
  Optimizers the fused kernel does not cover (SGD, amsgrad, fp32 parameters) take the same bucket pipeline with a
  reduce-to-owner kernel per owner run + a copy into the owner's persistent shard, and update at ``step()``.
* **ZeRO-3** — a tensor is resident on its owner only (the symmetric parameter buffer is sized for the
  largest owner share, not for the model).  ``fetch="push"`` (default): the owner ``multimem.st``-pushes a
  group of consecutively used tensors (≈ one layer, ≤ 16 MB) into a 4-slot symmetric staging ring on every
  rank, two groups ahead of the consumer, on the communication stream: each weight crosses NVLink once per
  pass.  ``fetch="peer"``: ``acquire`` hands the layer a tensor that *aliases the owner's memory over
  NVLink* and the tcgen05 GEMM's TMA producer streams the weight tiles peer-to-peer straight into shared
  memory (all-gather fused into the GEMM, nothing staged in HBM; M/128 re-reads of each tile).  Gradients
  are reduced to the owner by the same fused Adam kernel (no broadcast afterwards).

Synchronisation is device-side (flag pads), the whole step is CUDA-graph capturable, and nothing here
calls NCCL after construction.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from .. import ops
from ..nn.policy import CommPolicy
from . import symm

ALIGN = 64  # elements: every tensor starts on a 128-byte boundary of the flat buffers


def _pad(n: int) -> int:
    return (n + ALIGN - 1) // ALIGN * ALIGN


class NativePolicy(CommPolicy):
    is_native = True
    fused_rs = False          # EXPERIMENTAL GEMM -> reduce-scatter fusion (TDS_FUSED_RS=1), see __init__
    rs_names = frozenset()

    def __init__(self, mode: str, model: torch.nn.Module, *, table: Optional[Dict[str, int]] = None, group=None,
                 average: bool = False, bucket_bytes: int = 64 << 20, comm_blocks: int = 16,
                 grad_accumulation: bool = False, ring_slots: int = 3):
        self.mode = mode
        self.name = f"native-{mode}"
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.average = average
        self.scale = 1.0 / self.world if average else 1.0
        import os
        if self.world > 1:
            from .. import ops as _ops
            _ops.set_pdl(False)      # collectives run next to backward: no early-launched CTAs on the SM slots they need
        if os.environ.get("TDS_BUCKET_MB"):
            bucket_bytes = int(float(os.environ["TDS_BUCKET_MB"]) * (1 << 20))
        if os.environ.get("TDS_COMM_BLOCKS"):
            comm_blocks = int(os.environ["TDS_COMM_BLOCKS"])
        self.bucket_bytes = bucket_bytes
        # collectives that run UNDER backward take few CTAs (measured at 2 GPUs, ddp small: 8 / 16 / 32 / 64 blocks ->
        # 3.59 / 3.59 / 3.68 / 4.08 ms per step: every SM they sit on slows a single-wave GEMM); the ZeRO-3 parameter
        # push is on backward's critical path and keeps more requests in flight
        self.comm_blocks = comm_blocks
        self.push_blocks = int(os.environ.get("TDS_PUSH_BLOCKS", "32"))
        named = [(n, p) for n, p in model.named_parameters()]
        p0 = next(p for _, p in named if p.numel() > 0)
        self.device, self.dtype = p0.device, p0.dtype
        if self.dtype not in (torch.bfloat16, torch.float32):
            raise NotImplementedError("native backend trains bf16 (fp32 master weights in the optimizer) or fp32 parameters")
        if any(p.dtype != self.dtype for _, p in named):
            raise NotImplementedError("native backend needs one parameter dtype for the flat symmetric buffers")
        # fp32 models: the same NVLS kernels in their .f32 form (multimem.ld_reduce.add.v4.f32); the fused
        # reduce->Adam->multicast kernel is bf16-parameter only, fp32 takes reduce-to-owner + the multi-tensor Adam
        self.f32 = self.dtype == torch.float32
        self.esize = 4 if self.f32 else 2
        self.table = table or {n: 0 for n, _ in named}
        import os
        self.comm = symm.Comm(self.device, group)
        self.comm_stream = torch.cuda.Stream(self.device)
        # ZeRO bucket steps run on their own stream: under ZeRO-3 the communication stream carries the parameter pushes that
        # backward is waiting for, and a step kernel (Adam over a whole bucket + two cross-GPU barriers) must not delay them
        self.step_stream = torch.cuda.Stream(self.device) if mode == "zero3" else self.comm_stream

        # ---- layout ------------------------------------------------------------------------------
        self.names: List[str] = [n for n, _ in named]
        self.params: "OrderedDict[str, torch.nn.Parameter]" = OrderedDict(named)
        self.shape = {n: tuple(getattr(p, "_tds_shape", p.shape)) for n, p in named}
        self.numel = {n: int(torch.Size(self.shape[n]).numel()) for n in self.names}
        self._plan_gradient_layout(bucket_bytes, grad_accumulation, ring_slots)
        self.poff = {}                               # parameter buffer
        if mode == "zero3":
            share = [0] * self.world                 # owner-only layout: offsets inside the owner's region
            for n in self.names:
                r = self.table[n]
                self.poff[n] = share[r]
                share[r] += _pad(self.numel[n])
            self.ptotal = max(max(share), ALIGN)
        else:
            self.poff = dict(self.foff)
            self.ptotal = self.ftotal
        self.G = symm.alloc(self.gtotal * self.esize, self.device, group)
        self.P = symm.alloc(self.ptotal * self.esize, self.device, group)
        self.gflat = self.G.local.view(self.dtype)
        self.pflat = self.P.local.view(self.dtype)

        # ---- move parameters into the symmetric buffer (in place: the model keeps its Parameter objects) ----
        with torch.no_grad():
            for n, p in named:
                resident = mode != "zero3" or self.table[n] == self.rank
                view = self.pflat[self.poff[n]: self.poff[n] + self.numel[n]].view(self.shape[n])
                if resident:
                    if p.numel() == self.numel[n]:
                        view.copy_(p.data)
                    p.data = view
                else:
                    p.data = torch.empty(0, dtype=self.dtype, device=self.device)
                p._tds_shape = self.shape[n]
        self.gview = {n: self.gflat[self.goff[n]: self.goff[n] + self.numel[n]].view(self.shape[n]) for n in self.names}
        self._name_of = {id(p): n for n, p in named}
        torch.cuda.synchronize(self.device)
        self.comm.barrier()

        # ---- ZeRO-3 parameter fetch ----------------------------------------------------------------------------
        import os
        self.fetch = os.environ.get("TDS_ZERO3_FETCH", "push")     # "push": owner multicast + prefetch; "peer": GEMM pulls
        self.lookahead = int(os.environ.get("TDS_ZERO3_LOOKAHEAD", "2"))
        self.nslots = self.lookahead + 2
        self._seq, self._seq_frozen, self._pos, self._fetched = [], False, 0, {}
        self._groups, self._group_of = [], {}
        if mode == "zero3" and self.fetch == "push" and self.world > 1:
            self.slot_bytes = (max(_pad(v) for v in self.numel.values()) * self.esize + 4095) // 4096 * 4096
            self.S = symm.alloc(self.nslots * self.slot_bytes, self.device, group)
        # ---- EXPERIMENTAL (TDS_FUSED_RS=1, not yet exercised on multi-GPU hardware): GEMM -> reduce-scatter fusion ------
        # Every rank's dW GEMM epilogue adds its fp32 tile straight into the OWNER's reduction buffer over NVLink (TMA
        # reduce-add into peer memory, gemm_sm100.cu RED variant), so the gradient reduction of the Linear weights overlaps
        # backward tile by tile and the fused step reads an already-summed local buffer.  The buffer mirrors the layout of
        # the owner's optimizer state.  Embedding / LayerNorm / bias gradients keep the multimem path.
        self.fused_rs = (os.environ.get("TDS_FUSED_RS", "0") == "1" and mode != "ddp" and self.world > 1 and not self.f32)
        if self.fused_rs:
            from ..nn.modules import Linear
            lin = {id(m.weight) for m in model.modules() if isinstance(m, Linear)}
            self.rs_names = {n for n, p in named if id(p) in lin}
            self.rs_off, tot = {}, [0] * self.world
            for n in self.names:
                o = self.table[n]
                self.rs_off[n] = tot[o]
                tot[o] += _pad(self.numel[n])
            self.R = symm.alloc(max(max(tot), ALIGN) * 4, self.device, group)
            self.R.local.zero_()
            torch.cuda.synchronize(self.device)
            self.comm.barrier()
            rl = self.R.local.view(torch.float32)
            self._rs_local, self._rs_view = {}, {}
            for n in self.rs_names:
                o, off = self.table[n], self.rs_off[n]
                loc = rl[off: off + self.numel[n]].view(self.shape[n])
                view = loc if o == self.rank else self.R.peer(o, self.shape[n], torch.float32, off * 4)
                loc._tds_reduce = True
                view._tds_reduce = True
                self._rs_local[n], self._rs_view[n] = loc, view
        # row-sparse all-reduce of embedding gradients: validated (replicas stay bit-identical) but opt-in — at 2 GPUs it costs
        # 3.72 vs 3.60 ms/step: the dense 77 MB all-reduce of the last bucket already hides under the early Adam update, while
        # the sparse path adds four small latency-bound kernels (id copy, id all-gather, epoch bump, row reduce) to the tail
        self._sparse_emb = os.environ.get("TDS_SPARSE_EMB", "0") != "0" and mode == "ddp" and self.world > 1
        self._sparse_state = {}
        self._reset_round()
        self._accumulated = set()      # names holding un-synced micro-batch gradients
        self._opt_state = None
        self.stats = {"allreduce_launches": 0, "fused_steps": 0, "bytes": 0}
        # "exposed communication" meter: with the stub on, every collective is skipped / made rank-local,
        # so (step time) - (stubbed step time) is the communication the overlap did not hide.  Timing only: the
        # numerics of a stubbed step are meaningless.
        self.comm_stub = False
        ext = ops.ext()
        self._solo_ctx = ext.CommCtx([int(self.comm.flags.peer_ptrs[self.rank])], 0, 1, self.comm.error)
        self._solo_g = ext.SymmBuf([int(self.G.peer_ptrs[self.rank])], 0)
        self._solo_p = ext.SymmBuf([int(self.P.peer_ptrs[self.rank])], 0)

    def _plan_gradient_layout(self, bucket_bytes, grad_accumulation=False, ring_slots=3):
        """Buckets (reverse registration order) and the gradient-buffer layout: full (DDP, ZeRO-1, accumulation) or
        ring (ZeRO-2/3).  Pure host-side planning: no CUDA, covered by tests/test_native_logic_cpu.py."""
        import os
        # ---- bucketing: backward produces tensors roughly in reverse registration order ---------------------------------
        self.buckets: List[List[str]] = []
        cur, cur_bytes = [], 0
        for n in reversed(self.names):
            cur.append(n)
            cur_bytes += _pad(self.numel[n]) * self.esize
            if cur_bytes >= bucket_bytes:
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            self.buckets.append(cur)
        self.bucket_of = {n: i for i, b in enumerate(self.buckets) for n in b}
        # ---- gradient buffer layout ------------------------------------------------------------------------------------
        # full: every tensor, registration order (DDP, ZeRO-1, or gradient accumulation requested)
        # ring: ZeRO-2/3 — `nslots` bucket-sized slots recycled during backward (real gradient sharding: a rank never
        #       holds more than nslots buckets of full gradients + its own reduced shard)
        self.grad_accumulation = bool(grad_accumulation)
        # TDS_ZERO_OVERLAP=0: classic schedule (every bucket's step kernel after the last backward kernel) — needs the
        # full-size buffer, so it also switches the ring off
        self._zero_overlap = os.environ.get("TDS_ZERO_OVERLAP", "1") != "0"
        self.ring = self.mode in ("zero2", "zero3") and self.world > 1 and not self.grad_accumulation and \
            self._zero_overlap and os.environ.get("TDS_ZERO_RING", "1") != "0"
        self.foff, off = {}, 0                       # full layout: every tensor, registration order
        for n in self.names:
            self.foff[n] = off
            off += _pad(self.numel[n])
        self.ftotal = off
        self.goff = {}
        if self.ring:
            self.nslots_g = max(1, min(int(os.environ.get("TDS_RING_SLOTS", ring_slots)), len(self.buckets)))
            self.slot_elems = max(sum(_pad(self.numel[n]) for n in b) for b in self.buckets)
            for bi, b in enumerate(self.buckets):
                off = (bi % self.nslots_g) * self.slot_elems
                for n in b:
                    self.goff[n] = off
                    off += _pad(self.numel[n])
            self.gtotal = self.nslots_g * self.slot_elems
        else:
            self.goff = dict(self.foff)
            self.gtotal = self.ftotal

    def symmetric_bytes(self) -> int:
        """HBM held in symmetric allocations (not visible to torch's caching-allocator statistics)."""
        n = self.G.local.numel() + self.P.local.numel() + self.comm.flags.local.numel()
        if hasattr(self, "S"):
            n += self.S.local.numel()
        if getattr(self, "fused_rs", False):
            n += self.R.local.numel()
        return int(n)

    # ------------------------------------------------------------------------------------------ helpers
    overlap = None   # optim.overlap.StepOverlap bound to the comm stream (engine.TrainStep sets it for DDP)
    _opt = None      # sharded optimizer bound by its constructor (bind_optimizer): lets ZeRO buckets step inside backward

    def bind_optimizer(self, opt):
        """Called by the sharded optimizers' constructors: with the optimizer known, a completed ZeRO bucket can run its
        reduce -> Adam -> multicast kernel while backward is still producing the next one."""
        self._opt = opt

    def _reset_round(self):
        self._await_update = []
        self._bucket_event, self._launch_order = {}, []
        self._ready = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)
        self._synced_any = False
        self._step_opened = False
        self._zero_deferred = []          # complete ZeRO buckets waiting for one more grad_ready (their dX GEMMs)
        self._slot_waited = [False] * len(self.buckets)
        self._generic_round = False       # a bucket of this round went through the generic reduce-to-owner path
        self._rows = {}                   # DDP: embedding tables whose gradient is row-sparse this round -> token ids

    def _owner(self, name):
        return self.table[name]

    # ------------------------------------------------------------------------------------------ gradients
    def _wait_slot(self, b):
        """Ring layout: bucket b is about to be written into the slot bucket b - nslots used.  Every rank's step kernel
        of that older bucket must have finished (its trailing flag barrier is global), then the compute stream may go on."""
        old = b - self.nslots_g
        if old < 0 or self._slot_waited[b]:
            return
        if not self._launched[old]:
            if self._ready[old] != len(self.buckets[old]):
                raise RuntimeError(
                    f"native ZeRO gradient ring: bucket {b} starts before bucket {old} is complete — backward visits "
                    f"parameters too far out of registration order for {self.nslots_g} slots (raise TDS_RING_SLOTS or pass "
                    f"grad_accumulation=True for the full-size buffer)")
            self._flush_zero_buckets(upto=old)
        torch.cuda.current_stream(self.device).wait_event(self._bucket_event[old])
        self._slot_waited[b] = True

    def grad_out(self, param):
        n = self._name_of[id(param)]
        if self.fused_rs and n in self.rs_names:
            # the dW GEMM adds into the owner's fp32 buffer (ops.gemm sees `_tds_reduce`); stubbed runs stay rank-local
            return (self._rs_local[n] if self.comm_stub else self._rs_view[n]), False
        if self.ring:
            self._wait_slot(self.bucket_of[n])
        return self.gview[n], (n in self._accumulated)

    def grad_ready(self, param, grad, rows=None):
        n = self._name_of[id(param)]
        if self.fused_rs and n in self.rs_names:               # already on its way to the owner
            if getattr(param, "bwd_sync", False):
                param.bwd_sync = False
                self._synced_any = True
                self._zero_bucket_progress(n)
            return
        if grad.data_ptr() != self.gview[n].data_ptr():        # op ignored `out` (should not happen): copy in
            if n in self._accumulated:
                self.gview[n].add_(grad)
            else:
                self.gview[n].copy_(grad)
        fresh = n not in self._accumulated                     # no earlier micro-batch accumulated into this buffer
        self._accumulated.add(n)
        owner_like = self.mode in ("ddp", "zero1") or self._owner(n) == self.rank
        if owner_like and param.numel() > 0 and not self.ring:
            param.grad = self.gview[n]
        if not getattr(param, "bwd_sync", False):
            if self.ring:
                raise NotImplementedError("gradient accumulation (a backward without grad sync) needs the full-size gradient "
                                          "buffer: wrap with Zero2/Zero3(..., grad_accumulation=True)")
            return
        param.bwd_sync = False
        self._synced_any = True
        if self.world == 1:
            return
        if self.mode == "ddp":
            b = self.bucket_of[n]
            self._ready[b] += 1
            if rows is not None and self._sparse_ok(n, rows, fresh):
                self._rows[n] = rows                   # only these rows of the table's gradient are non-zero
            # the all-reduce only touches the gradient buffer, so a bucket can go the moment its last dW is enqueued;
            # it then runs on the comm stream underneath the remaining dX/dW GEMMs of backward
            # optimizer-in-backward: buckets whose all-reduce was queued at an EARLIER grad_ready can be updated now
            # (their layers' dX GEMMs are enqueued); the update runs on the comm stream right behind the all-reduce
            if self.overlap is not None:
                for pb in self._await_update:
                    self.overlap._launch(list(self.buckets[pb]), overlapped=True)
                self._await_update = []
            if self._ready[b] == len(self.buckets[b]) and not self._launched[b]:
                self._launch_bucket(b)
                self._await_update.append(b)
        else:
            self._zero_bucket_progress(n)

    # ---- ZeRO-1/2/3: a completed bucket goes to the communication stream while backward continues ---------------------
    def _zero_bucket_progress(self, n):
        b = self.bucket_of[n]
        # buckets completed at an EARLIER grad_ready: the dX GEMMs that still read their parameters are enqueued by now,
        # so the step kernel (which rewrites those parameters on every rank) may be ordered behind the compute stream
        if self._zero_deferred:
            for pb in self._zero_deferred:
                self._launch_zero_bucket(pb)
            self._zero_deferred = []
        self._ready[b] += 1
        if self._zero_overlap and self._ready[b] == len(self.buckets[b]) and not self._launched[b]:
            self._zero_deferred.append(b)

    def _flush_zero_buckets(self, upto=None):
        """Launch every complete bucket that has not gone yet (all of them at the end of backward)."""
        for b in list(self._zero_deferred):
            if upto is None or b <= upto:
                self._launch_zero_bucket(b)
        self._zero_deferred = [b for b in self._zero_deferred if not self._launched[b]]
        if upto is None:
            for b, names in enumerate(self.buckets):
                if not self._launched[b] and self._ready[b] > 0:
                    self._launch_zero_bucket(b)

    def _fusable(self, opt) -> bool:
        from ..optim.adamw import AdamW
        return opt is not None and isinstance(opt, AdamW) and not opt.amsgrad and not self.f32 and self.world > 1

    def _launch_zero_bucket(self, b):
        if self._launched[b]:
            return
        names = self.buckets[b]
        opt = self._opt
        fused = self._fusable(opt) and not self._generic_round
        if not fused:
            self._generic_round = True
        cur = torch.cuda.current_stream(self.device)
        if fused:
            st = self._ensure_opt_state(opt)
            if not self._step_opened:                  # this step's counter, once, on the compute stream before the fork
                opt.step_count += 1
                self._step_opened = True
            step_dev = opt._device_step(self.device)
        self.step_stream.wait_stream(cur)
        with torch.cuda.stream(self.step_stream):
            if fused:
                self._fused_bucket_step(opt, st, b, step_dev)
            else:
                self._reduce_bucket_to_owners(b)
            ev = torch.cuda.Event()
            ev.record(self.step_stream)
        self._bucket_event[b] = ev
        self._launch_order.append(b)
        self._launched[b] = True
        self.stats["bytes"] += sum(_pad(self.numel[n]) for n in names) * self.esize

    def _bucket_runs(self, b):
        """Owner runs of bucket b in gradient-buffer order: [(owner, [names...])] with contiguous gradient offsets."""
        runs = []
        for n in sorted(self.buckets[b], key=lambda k: self.goff[k]):
            o = self._owner(n)
            if runs and runs[-1][0] == o and self.goff[runs[-1][1][-1]] + _pad(self.numel[runs[-1][1][-1]]) == self.goff[n]:
                runs[-1][1].append(n)
            else:
                runs.append((o, [n]))
        return runs

    def _reduce_bucket_to_owners(self, b):
        """Generic optimizers (SGD, amsgrad, fp32 parameters): one reduce-to-owner kernel per owner run of the bucket
        (instead of one per tensor), then — ring layout — a copy of the owner's reduced range into its persistent shard."""
        for owner, names in self._bucket_runs(b):
            lo = self.goff[names[0]]
            n_el = sum(_pad(self.numel[k]) for k in names)
            if not self.comm_stub:
                self.comm.reduce_to(self.G, lo, n_el, owner, f32=self.f32, scale=self.scale, blocks=self.comm_blocks, channel=1)
            if owner == self.rank:
                for k in names:
                    p = self.params[k]
                    if self.ring:
                        dst = self._own_grad(k)
                        dst.copy_(self.gview[k])
                        p.grad = dst
                    elif p.numel() > 0:
                        p.grad = self.gview[k]

    def _own_grad(self, name):
        """Ring layout: persistent home of an OWNED tensor's reduced gradient (the ring slot is recycled)."""
        if getattr(self, "_gown", None) is None:
            owned = [k for k in self.names if self._owner(k) == self.rank]
            self._gown_off, off = {}, 0
            for k in owned:
                self._gown_off[k] = off
                off += _pad(self.numel[k])
            self._gown = torch.zeros(max(off, ALIGN), dtype=self.dtype, device=self.device)
        o = self._gown_off[name]
        return self._gown[o: o + self.numel[name]].view(self.shape[name])

    def _launch_complete_buckets(self, flush=False):
        for b, names in enumerate(self.buckets):
            if self._launched[b]:
                continue
            if self._ready[b] == len(names) or (flush and self._ready[b] > 0):
                # the flush at the end of backward has the GPU to itself: use every comm block the flag pad allows
                self._launch_bucket(b, blocks=ops.ext().COMM_MAX_BLOCKS if flush else None)

    # ---- DDP: row-sparse all-reduce of embedding gradients ----------------------------------------------------------
    def _sparse_ok(self, n, rows, fresh) -> bool:
        """Reduce only the touched rows of table `n`?  bf16, plain SUM, no accumulated micro-batches, and the rows of all
        ranks together cover at most a quarter of the table (wte: 8 x 1024 of 50304 rows; wpe: every row -> dense)."""
        if not self._sparse_emb or self.f32 or self.scale != 1.0 or not fresh or self.comm_stub:
            return False
        shape = self.shape[n]
        return (len(shape) == 2 and shape[1] % 8 == 0 and rows.numel() % 2 == 0 and rows.dtype == torch.int64
                and self.world * rows.numel() * 4 <= shape[0])

    def _sparse_allreduce(self, n, rows, blocks):
        """On the communication stream: all-gather the token ids, then switch-reduce + multicast only the touched rows."""
        ext = ops.ext()
        ntok = rows.numel()
        st = self._sparse_state.get(n)
        if st is None or st["ntok"] != ntok:
            ids = symm.alloc(self.world * ntok * 8, self.device, self.group)
            st = dict(ntok=ntok, ids=ids, flat=ids.local.view(torch.int64),
                      epoch_of_row=torch.zeros(self.shape[n][0], dtype=torch.int32, device=self.device),
                      epoch=torch.zeros(1, dtype=torch.int32, device=self.device))
            self._sparse_state[n] = st
        st["flat"][self.rank * ntok: (self.rank + 1) * ntok].copy_(rows.reshape(-1))
        ext.comm_allgather_slots(self.comm.ctx, st["ids"].buf, 0, ntok * 8, 4, 0)
        ext.step_increment(st["epoch"])
        ext.comm_allreduce_rows(self.comm.ctx, self.G.buf, self.goff[n] * self.esize, self.shape[n][1] * self.esize,
                                st["flat"][: self.world * ntok], self.shape[n][0], st["epoch_of_row"], st["epoch"],
                                int(blocks), 0)
        ops.count_launch(3)
        self.stats["sparse_allreduce_launches"] = self.stats.get("sparse_allreduce_launches", 0) + 1
        self.stats["bytes"] += self.world * ntok * (8 + self.shape[n][1] * self.esize)

    def _launch_bucket(self, b, blocks=None):
        names = self.buckets[b]
        # embedding tables with a row list leave the dense range when they sit at one end of it (wte is the first tensor of
        # the flat layout, i.e. the low end of the last bucket)
        dense = sorted(names, key=lambda k: self.goff[k])
        sparse = []
        while dense and dense[0] in self._rows:
            sparse.append(dense.pop(0))
        while dense and dense[-1] in self._rows:
            sparse.append(dense.pop())
        cur = torch.cuda.current_stream(self.device)
        self.comm_stream.wait_stream(cur)
        with torch.cuda.stream(self.comm_stream):
            if not self.comm_stub:
                if dense:
                    lo = min(self.goff[n] for n in dense)
                    hi = max(self.goff[n] + _pad(self.numel[n]) for n in dense)
                    self.comm.allreduce(self.G, lo, hi - lo, f32=self.f32, scale=self.scale, blocks=blocks or self.comm_blocks,
                                        channel=0)
                    self.stats["bytes"] += (hi - lo) * self.esize
                for n in sparse:
                    self._sparse_allreduce(n, self._rows[n], blocks or self.comm_blocks)
            ev = torch.cuda.Event()
            ev.record(self.comm_stream)
        self._bucket_event[b] = ev
        self._launch_order.append(b)
        self._launched[b] = True
        self.stats["allreduce_launches"] += 1

    def flush_async(self):
        """DDP: queue the all-reduce of the last (still open) buckets WITHOUT joining the communication stream and
        return the names whose gradients are therefore not final yet.  The optimizer updates every other tensor
        first — ~0.5 ms of HBM-bound Adam that hides the final all-reduce (the embedding gradients are produced by
        the very last backward kernel, so nothing else can overlap it) — then calls :meth:`join`."""
        if not (self.mode == "ddp" and self.world > 1 and self._synced_any):
            self.finish()
            return set()
        self._launch_complete_buckets(flush=True)
        if not self._launch_order:
            return set()
        # the most recently queued all-reduce (it carries the embedding gradients, produced by the last backward
        # kernel) is treated as in flight; everything queued before it is ordered in front of the early update by
        # waiting on the event recorded behind the previous bucket (comm stream is in-order)
        late_b = self._launch_order[-1]
        if len(self._launch_order) > 1:
            torch.cuda.current_stream(self.device).wait_event(self._bucket_event[self._launch_order[-2]])
        self._join_pending = True
        return set(self.buckets[late_b])

    def join(self):
        if getattr(self, "_join_pending", False):
            torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)
            self._join_pending = False
            self._accumulated.clear()
            self._reset_round()

    def finish(self):
        """Join the communication stream into the compute stream (no host synchronisation)."""
        if self.mode == "ddp" and self.world > 1 and self._synced_any:
            self._launch_complete_buckets(flush=True)
            torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)
            self._join_pending = False
            self._accumulated.clear()
            self._reset_round()
        elif self.mode != "ddp" and self.world > 1 and self._synced_any and not self._step_opened:
            # generic optimizer: every bucket is reduced onto its owners (most of them already were, during backward)
            self._flush_zero_buckets()
            torch.cuda.current_stream(self.device).wait_stream(self.step_stream)
            self._accumulated.clear()
            self._reset_round()
            self._end_round_zero3()

    # ------------------------------------------------------------------------------------------ ZeRO-3 parameters
    def acquire(self, param, *, backward=False, sparse=False):
        if self.mode != "zero3" or self.world == 1:
            return param
        n = self._name_of[id(param)]
        owner = self._owner(n)
        if sparse and owner != self.rank:
            # embedding gather: <= ntokens rows of a [V, D] table (1.5 MB of the 77 MB wte for GPT-2 small) — the gather
            # kernel reads them straight from the owner's memory over NVLink instead of waiting for a push of the whole
            # table at the very start of forward, where nothing can hide it
            return self.P.peer(owner, self.shape[n], self.dtype, self.poff[n] * self.esize)
        if sparse:
            return param.data
        if self.fetch == "peer":
            if owner == self.rank:
                return param.data
            # alias of the OWNER's memory (NVLink peer mapping): the consuming GEMM's TMA producer pulls the weight
            # tile by tile straight into shared memory (all-gather fused into the GEMM; M/128 x NVLink re-reads)
            t = self.P.peer(owner, self.shape[n], self.dtype, self.poff[n] * self.esize)
            t._tds_remote = True
            return t
        # ---- owner-push mode: the owner multicasts a GROUP of consecutively used tensors (a contiguous byte range of
        # its region, typically one transformer layer) into a staging slot of every rank, up to `lookahead` groups
        # ahead of the consumer, on the communication stream --------------------------------------------------------
        pos = self._pos
        self._pos += 1
        if self._seq_frozen and (pos >= len(self._seq) or self._seq[pos] != n):
            self._seq_frozen, self._seq, pos = False, [], 0       # use order changed: re-record from here
            self._pos = 1
            self._fetched.clear()
            self._groups, self._group_of = [], {}
        if not self._seq_frozen:
            # recording step: one group per use (no lookahead yet)
            self._seq.append(n)
            self._groups.append(dict(first=pos, owner=owner, lo=self.poff[n], hi=self.poff[n] + _pad(self.numel[n])))
            self._group_of[pos] = len(self._groups) - 1
        g = self._group_of[pos]
        if g not in self._fetched:
            self._launch_fetch(g)
        if self._seq_frozen:
            for la in range(1, self.lookahead + 1):
                if g + la < len(self._groups) and (g + la) not in self._fetched:
                    self._launch_fetch(g + la)
        torch.cuda.current_stream(self.device).wait_event(self._fetched[g])
        if owner == self.rank:
            return param.data
        grp = self._groups[g]
        off = (g % self.nslots) * self.slot_bytes + (self.poff[n] - grp["lo"]) * self.esize
        return self.S.local[off: off + self.numel[n] * self.esize].view(self.dtype).view(self.shape[n])

    def _launch_fetch(self, g):
        grp = self._groups[g]
        owner, slot = grp["owner"], g % self.nslots
        nbytes = (grp["hi"] - grp["lo"]) * self.esize
        src = self.pflat.data_ptr() + grp["lo"] * self.esize if owner == self.rank else 0
        # everything enqueued so far on the compute stream precedes the push: the slot's previous readers and the
        # optimizer update of the source range are therefore complete when the owner starts writing
        self.comm_stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.comm_stream):
            if not self.comm_stub:
                ops.ext().comm_push(self.comm.ctx, int(src), self.S.buf, slot * self.slot_bytes, nbytes, owner,
                                    self.push_blocks, 2)
                ops.count_launch()
            ev = torch.cuda.Event()
            ev.record(self.comm_stream)
        self._fetched[g] = ev
        self.stats["bytes"] += nbytes

    def _build_groups(self):
        """Merge the recorded use sequence into fetch groups: consecutive uses owned by the same rank whose byte ranges
        (in the owner's region) stay within one staging slot and leave gaps of at most `gap` elements."""
        # elements: <= 16 MB per push keeps the pipeline fine-grained
        gap, cap = 1 << 19, min(self.slot_bytes // self.esize, (16 << 20) // self.esize)
        groups, group_of = [], {}
        for pos, n in enumerate(self._seq):
            lo, hi, owner = self.poff[n], self.poff[n] + _pad(self.numel[n]), self._owner(n)
            if groups:
                cur = groups[-1]
                nlo, nhi = min(cur["lo"], lo), max(cur["hi"], hi)
                near = lo <= cur["hi"] + gap and hi + gap >= cur["lo"]
                # a group may be live while the next `lookahead` groups are being written: keep at most 2 uses of
                # slack by never letting one group span more than a slot
                if cur["owner"] == owner and near and nhi - nlo <= cap:
                    cur["lo"], cur["hi"] = nlo, nhi
                    group_of[pos] = len(groups) - 1
                    continue
            groups.append(dict(first=pos, owner=owner, lo=lo, hi=hi))
            group_of[pos] = len(groups) - 1
        self._groups, self._group_of = groups, group_of

    def _end_round_zero3(self):
        if self.mode == "zero3" and self.fetch == "push":
            if not self._seq_frozen and self._seq:
                self._seq_frozen = True
                self._build_groups()
            self._pos = 0
            self._fetched.clear()

    def release(self, param, full):
        return

    # ------------------------------------------------------------------------------------------ generic optimizer path
    def broadcast_params(self):
        """ZeRO-1/2 without the fused step (SGD, amsgrad, fp32 parameters): every owner multicasts its freshly updated
        tensors; consecutive tensors of one owner travel as one contiguous byte range (one kernel each).  Replaces the
        reference's per-tensor blocking ``dist.broadcast`` (zero1/optim.py:20-34)."""
        if self.mode not in ("zero1", "zero2") or self.world == 1:
            return
        if getattr(self, "_bcast_ranges", None) is None:
            ranges = []
            for n in self.names:
                lo, hi, owner = self.poff[n], self.poff[n] + _pad(self.numel[n]), self._owner(n)
                if ranges and ranges[-1][2] == owner and ranges[-1][1] == lo:
                    ranges[-1][1] = hi
                else:
                    ranges.append([lo, hi, owner])
            self._bcast_ranges = ranges
        if self.comm_stub:
            return
        for lo, hi, owner in self._bcast_ranges:
            self.comm.broadcast(self.P, lo * self.esize, (hi - lo) * self.esize, owner, blocks=self.comm_blocks, channel=0)

    # ------------------------------------------------------------------------------------------ fused optimizer step
    def owns_optimizer_state(self, opt) -> bool:
        return self.mode != "ddp" and self._fusable(opt)

    def _ensure_opt_state(self, opt):
        if self._opt_state is not None:
            return self._opt_state
        owned = [n for n in self.names if self.mode == "ddp" or self._owner(n) == self.rank]
        soff, off = {}, 0
        for n in owned:
            soff[n] = off
            off += _pad(self.numel[n])
        total = max(off, ALIGN)
        master = torch.zeros(total, dtype=torch.float32, device=self.device)
        m = torch.zeros(total, dtype=torch.float32, device=self.device)
        v = torch.zeros(total, dtype=torch.float32, device=self.device)
        for n in owned:
            src = self.pflat[self.poff[n]: self.poff[n] + self.numel[n]]
            master[soff[n]: soff[n] + self.numel[n]].copy_(src.float())
            # expose the compact state through the optimizer's public per-name dict (checkpointing, tests)
            sl = slice(soff[n], soff[n] + self.numel[n])
            opt.state[n] = {"exp_avg": m[sl].view(self.shape[n]), "exp_avg_sq": v[sl].view(self.shape[n]),
                            "master": master[sl].view(self.shape[n])}
        max_ranges = int(ops.ext().COMM_MAX_RANGES)
        owned_set = set(owned)
        bucket_ranges, bucket_min = [], []
        for names in self.buckets:
            mine = sorted((n for n in names if n in owned_set), key=lambda k: self.goff[k])
            rs = [[self.goff[n], _pad(self.numel[n]), soff[n], self.poff[n]] +
                  ([int(n in self.rs_names)] if self.fused_rs else []) for n in mine]
            bucket_ranges.append(rs)
            # every rank must launch the same number of (barrier-carrying) kernels per bucket: the rank owning most decides
            per_rank = [sum(1 for n in names if self._owner(n) == r) for r in range(self.world)]
            bucket_min.append(max(1, max((c + max_ranges - 1) // max_ranges for c in per_rank)))
        self._opt_state = dict(master=master, m=m, v=v, owned=owned, bucket_ranges=bucket_ranges, bucket_min=bucket_min)
        return self._opt_state

    def _fused_bucket_step(self, opt, st, b, step_dev):
        """reduce -> scale -> Adam -> (multicast) for the tensors of bucket b this rank owns; every rank launches (the
        kernel's flag barriers are collective) whatever it owns."""
        ext = ops.ext()
        ctx, gbuf, pbuf = (self._solo_ctx, self._solo_g, self._solo_p) if self.comm_stub else (self.comm.ctx, self.G.buf, self.P.buf)
        hyper = (float(opt.lr), float(opt.beta1), float(opt.beta2), float(opt.eps), float(opt.weight_decay), step_dev,
                 bool(opt.decoupled), bool(opt.maximize), float(opt.grad_scale * self.scale), self.mode != "zero3", 1,
                 int(st["bucket_min"][b]))
        if self.fused_rs:
            if self.comm_stub and not hasattr(self, "_solo_r"):
                self._solo_r = ext.SymmBuf([int(self.R.peer_ptrs[self.rank])], 0)
            rbuf = self._solo_r if self.comm_stub else self.R.buf
            launches = ext.comm_zero_fused_adam_rs(ctx, gbuf, pbuf, rbuf, st["bucket_ranges"][b], st["master"], st["m"], st["v"],
                                                   *hyper)
        else:
            launches = ext.comm_zero_fused_adam(ctx, gbuf, pbuf, st["bucket_ranges"][b], st["master"], st["m"], st["v"], *hyper)
        ops.count_launch(int(launches))
        self.stats["fused_steps"] += 1

    def fused_optimizer_step(self, opt) -> bool:
        """ZeRO-1/2/3 + Adam: called from ``optimizer.step()``.  Most buckets already went through their fused
        reduce -> Adam -> (multicast) kernel during backward; this launches the rest and joins the communication stream.
        Returns False when this policy/optimizer pair must take the generic path."""
        if self.mode == "ddp" or not self._fusable(opt):
            return False
        if not self._synced_any or self._generic_round:
            return False
        self._opt = opt
        self._flush_zero_buckets()
        # (the ZeRO-3 push stream needs no join here: every push was consumed through its event by the layer that used it)
        torch.cuda.current_stream(self.device).wait_stream(self.step_stream)
        for p in self.params.values():
            p.grad = None
        self._accumulated.clear()
        self._reset_round()
        self._end_round_zero3()
        return True
