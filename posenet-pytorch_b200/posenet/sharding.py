"""Multi-GPU sharding of the hot path: one process per GPU, images split across ranks, weights
replicated, NO collective on the data path.  The reference has no multi-GPU code;
images are independent, so the only exchange is collecting the fixed-size pose records
(``86 * P`` float64 values per image) after a batch -- an all-gather over NCCL/NVLink on GPUs, gloo
in the CPU tests.
"""
import torch
import torch.distributed as dist

from posenet.constants import NUM_KEYPOINTS


def shard_bounds(n_items, rank, world):
    """Contiguous block of ``n_items`` owned by ``rank``: sizes differ by at most one, earlier ranks
    take the remainder.  Returns ``(begin, end)``."""
    assert 0 <= rank < world and n_items >= 0
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def pack_pose_records(pose_scores, keypoint_scores, keypoint_coords, pose_offsets):
    """[n,P], [n,P,17], [n,P,17,2], [n,P,17,2] -> one [n, P*86] float64 tensor (one row per image)."""
    n, P = pose_scores.shape
    return torch.cat([pose_scores.reshape(n, P), keypoint_scores.reshape(n, P * NUM_KEYPOINTS),
                      keypoint_coords.reshape(n, P * NUM_KEYPOINTS * 2),
                      pose_offsets.reshape(n, P * NUM_KEYPOINTS * 2)], dim=1).contiguous()


def unpack_pose_records(rows, max_pose_detections):
    n, P, K = rows.shape[0], int(max_pose_detections), NUM_KEYPOINTS
    assert rows.shape[1] == P * (1 + 5 * K), "row width %d does not match P=%d" % (rows.shape[1], P)
    a, b, c = P, P * (1 + K), P * (1 + 3 * K)
    return (rows[:, :a].reshape(n, P), rows[:, a:b].reshape(n, P, K), rows[:, b:c].reshape(n, P, K, 2),
            rows[:, c:].reshape(n, P, K, 2))


def gather_pose_records(pose_scores, keypoint_scores, keypoint_coords, pose_offsets, n_total=None, group=None):
    """All-gather every rank's pose records into image order.

    Each rank passes the records of its ``shard_bounds`` block (shards may differ by one image; they
    are padded to the largest shard for the fixed-size collective and trimmed afterwards).  Returns the
    same 4-tuple for all ``n_total`` images on every rank.  Without an initialised process group this
    is the identity."""
    P = pose_scores.shape[1]
    rows = pack_pose_records(pose_scores, keypoint_scores, keypoint_coords, pose_offsets)
    if not (dist.is_available() and dist.is_initialized()):
        return unpack_pose_records(rows, P)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if n_total is None:
        cnt = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
        dist.all_reduce(cnt, group=group)
        n_total = int(cnt.item())
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    assert rows.shape[0] == sizes[rank][1] - sizes[rank][0], "rank %d holds %d images, its shard is %s" % (
        rank, rows.shape[0], sizes[rank])
    width = max(e - b for b, e in sizes)
    padded = rows.new_zeros((width, rows.shape[1]))
    padded[:rows.shape[0]] = rows
    out = rows.new_empty((world * width, rows.shape[1]))
    dist.all_gather_into_tensor(out, padded, group=group)
    parts = [out[r * width:r * width + (e - b)] for r, (b, e) in enumerate(sizes)]
    return unpack_pose_records(torch.cat(parts, dim=0), P)


def infer_sharded(model, images_u8, output_stride=None, group=None, **decode_kw):
    """images_u8: the FULL uint8 [N,H,W,3] batch (host or device), identical on every rank.  Each rank
    runs preprocess+backbone+decode on its block on its own GPU and the records are gathered, so every
    rank returns the poses of all N images (bit-identical to a single-GPU run: images are independent
    and every kernel is deterministic)."""
    from posenet.decode_multi import decode_multiple_poses_batch
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    b, e = shard_bounds(images_u8.shape[0], rank, world)
    mine = images_u8[b:e].to(model._device(), non_blocking=True)
    heads = model.forward_u8(mine)
    ps, ks, kc, ko, _ = decode_multiple_poses_batch(*heads, output_stride=output_stride or model.output_stride, **decode_kw)
    return gather_pose_records(ps, ks, kc, ko, n_total=images_u8.shape[0], group=group)
