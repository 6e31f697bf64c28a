"""Module-level knobs of the engine (the reference's constructors are left untouched: SURVEY.md §5)."""
import os


def _flag(name, default):
    v = os.environ.get(name)
    if v is None:
        return default
    return v.strip().lower() not in ("0", "false", "no", "off", "")


# Route the embedding tables of CTRTrainer to the fused row-wise (lazy) optimiser that mirrors the
# configured torch optimiser on TOUCHED ROWS ONLY (SURVEY.md §7 hard part 2).  Off by default: the
# default keeps the reference's exact dense-optimiser semantics (dense ``weight.grad`` + torch.optim).
rowwise_optimizer = _flag("RECHUB_B200_ROWWISE_OPT", False)

# Capture CTRTrainer's training step into a CUDA graph (static shapes only; ragged last batches run eagerly).
cuda_graph = _flag("RECHUB_B200_CUDA_GRAPH", False)

# Also capture the sharded (multi-GPU) step, NCCL collectives included, into the graph.
dist_cuda_graph = _flag("RECHUB_B200_DIST_CUDA_GRAPH", True)

# Sharded front end: exchange ids / rows / row-gradients with the engine's own kernels over NVLink peer memory
# (torch symmetric memory) instead of NCCL all-to-alls.
p2p_exchange = _flag("RECHUB_B200_P2P", True)

# Peer-memory exchange details.  Field-major id buffers: contiguous NVLink stores from rh_ids_scatter and unit-stride ids for the
# owner's gather (off = sample-major, the first layout).  Direct gradients: the sample's GPU REDs each row gradient straight into
# the OWNER's persistent gradient buffer at the row id (the buffers live in one symmetric-memory pool per rank); off = RED into a
# staging buffer on the owner + an owner-side scatter-add pass.
p2p_field_major_ids = _flag("RECHUB_B200_P2P_FIELD_MAJOR", True)
p2p_direct_grads = _flag("RECHUB_B200_P2P_DIRECT_GRADS", True)
# Issue the post-backward barrier (all row-gradient REDs landed) AFTER launching the dense all-reduce, so that it overlaps it.
# The barrier and the all-reduce are then concurrent branches of the captured graph and BOTH wait on peers: correct only while
# every rank's executor orders the two branches the same way.  Measured on 2 GPUs only -> off by default.
p2p_defer_barrier = _flag("RECHUB_B200_P2P_DEFER_BARRIER", False)

# Sharded step: all-reduce the replicated parameters' gradients over NVLink peer memory inside the engine's own kernels, fused with
# the dense optimiser update (rh_dense_pack_signal + rh_dense_reduce_update) instead of NCCL's all-reduce + rh_dense_update.
p2p_allreduce = _flag("RECHUB_B200_P2P_ALLREDUCE", True)

# ... and order the exchange's hand-overs with the engine's own flag barrier (rh_peer_barrier, one warp) instead of the symmetric-memory
# library's barrier kernel.
p2p_own_barrier = _flag("RECHUB_B200_P2P_OWN_BARRIER", True)
# With the peer-memory reduction the post-backward barrier is a WAIT on the gradient-publication flags (rh_peer_wait): a rank publishes
# behind its backward kernels and a system fence, so no separate signal round is needed.
p2p_fold_barrier = _flag("RECHUB_B200_P2P_FOLD_BARRIER", True)
# The forward exchange's two barriers (+ the owner's id snapshot copy) folded into its three launches: rh_ids_scatter_signal publishes
# "my ids have landed" when its last CTA is done, the owner's gather (rh_fields_fwd_sync) waits for every rank's ids inside the kernel,
# keeps a local copy of the ids it reads and publishes "my rows have landed", the sample side's unpack launch waits for the owners'
# rows inside the kernel.  6 graph nodes -> 3.  Off: rh_ids_scatter | barrier | copy | rh_fields_fwd_p2p | barrier | rh_fields_fwd.
p2p_fused_sync = _flag("RECHUB_B200_P2P_FUSED_SYNC", True)

# Check the device-side out-of-range-id flag after every forward (one D2H sync per step).  When off the
# flag is checked at the trainer's existing sync points (``loss.item()``) and by ``check_errors()``.
eager_bounds_check = _flag("RECHUB_B200_EAGER_BOUNDS_CHECK", False)

# Tower GEMMs on the tcgen05 tensor cores with the fp32-accurate 3xTF32 kernel (rh_gemm_tf32x3); off = cuBLAS fp32 (torch.mm).
tensor_core_gemm = _flag("RECHUB_B200_TC_GEMM", True)

# Graph-replayed training loop: ship host batches over PCIe on a copy stream while the previous step computes
# (staging buffers + D2D into the graph's static inputs) ...
pipelined_inputs = _flag("RECHUB_B200_PIPELINED_INPUTS", True)

# ... and read each step's loss / out-of-range-id flag one step late (async D2H into pinned slots), so the training
# loop never drains the GPU.  Epoch statistics are unchanged; an IndexError surfaces one batch later than in the reference.
lagged_loss = _flag("RECHUB_B200_LAGGED_LOSS", True)

# Tower backward: run a layer's weight-gradient GEMM on a second stream next to its input-gradient GEMM (they only share
# d_h, and each fills about half of the 148 SMs at batch 4096).
concurrent_tower_bwd = _flag("RECHUB_B200_CONCURRENT_BWD", True)

# Hybrid optimiser: the row-wise table update and the dense tower update on two streams (they share only the step counter).
concurrent_optimizers = _flag("RECHUB_B200_CONCURRENT_OPT", False)  # measured: no gain at batch 4096 (0.2258 vs 0.2253 ms)

# Output head (Linear(K,1) + side terms + sigmoid) as one launch each way (rh_head_fwd/bwd) instead of ~10 library launches.
fused_head = _flag("RECHUB_B200_FUSED_HEAD", True)

# The same head for the other ranking models (DCN / DCNv2: LR over [cross | deep]; WideDeep: wide term + deep head; DIN: the final
# tower).  Off until its GPU validation: run the GPU suite with RECHUB_B200_FUSED_HEAD_ALL=1.
fused_head_all = _flag("RECHUB_B200_FUSED_HEAD_ALL", False)

# Pipelined loop: prefetch into L2 the table / gradient / optimiser rows the NEXT batch will touch (rh_fields_prefetch on the copy
# stream, one step ahead).  Written after the round's last GPU session -> off until measured.
next_batch_prefetch = _flag("RECHUB_B200_NEXT_BATCH_PREFETCH", False)

# Training-mode BatchNorm + activation + dropout of a tower layer as ONE launch each way (rh_bn_act_fused_fwd/_bwd: rows in
# registers across a grid barrier) instead of rh_colstats + rh_bn_act_fwd and the two passes of rh_bn_act_bwd ...
fused_bn = _flag("RECHUB_B200_FUSED_BN", True)
# ... and the last hidden layer's fused launch also does the tower's output layer + side terms + sigmoid (DeepFM / DIN heads).
fused_bn_head = _flag("RECHUB_B200_FUSED_BN_HEAD", True)

# Programmatic dependent launch of the hot-path kernels (rh_set_pdl): a kernel's scheduling and memory-free prologue overlap the drain
# of its predecessor; captured as programmatic edges by CUDA graphs.  Off until measured.
pdl = int(os.environ.get("RECHUB_B200_PDL", "0"), 0)  # mask of kernel families launched early (rh_set_pdl); 0xff = all

# cudaLimitMaxL2FetchGranularity the engine sets on every device it touches (bytes; 0 = leave the process default).  Random 64-byte
# table rows are the dominant DRAM access: measured DRAM reads per 106 k-row gather: 14.4 MB at 128, 7.7 MB (= algorithmic) at 64 / 32.
l2_fetch_granularity = int(os.environ.get("RECHUB_B200_L2_FETCH_GRANULARITY", "32"))

# The optimiser's step-counter / bias-correction launch (rh_opt_advance) is issued at the START of the step on the side stream
# (RowwiseOptimizer.advance_early) instead of between the scatter-add and the row-wise update.
# Measured (profiles/README.md, r02c): 0.1625 ms with, 0.1617 ms without — no gain; off.
early_opt_advance = _flag("RECHUB_B200_EARLY_OPT_ADVANCE", False)

# The split-K targets of the tower's weight-gradient GEMMs are zeroed during the FORWARD on the idle side stream (ops._prezero_dw).
# Measured (profiles/README.md, r02d): 0.1638 ms with, 0.1628 ms without — no gain; off.
prezero_dw = _flag("RECHUB_B200_PREZERO_DW", False)

# Set by the graph runner while inputs live in static buffers that the next batch overwrites.
static_inputs = False

# Output-tile width of the tower GEMM (rh_gemm_tile_n): 0 = the library chooses per problem (128 x 64 tiles when 128 x 128 would leave
# more than half of the SMs idle), 64 / 128 force one (A/B runs).
gemm_tile_n = int(os.environ.get("RECHUB_B200_GEMM_TILE_N", "0"))

# Preferred shared-memory carveout for every kernel of the library (rh_set_smem_carveout): -1 = the driver's choice.  Experiment switch.
smem_carveout = int(os.environ.get("RECHUB_B200_SMEM_CARVEOUT", "-1"))
