"""Frame-batch sharding across GPUs: stereo pairs are independent, so rank r simply owns a contiguous range of
frames (frame i -> rank floor(i * world / n)); there is no data-path collective."""


def shard_range(n_items, rank, world):
    """[start, stop) of the items owned by `rank`; ranges are contiguous, disjoint and cover [0, n_items)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    start = -(-rank * n_items // world)          # ceil(rank * n / world): first i with floor(i*world/n) == rank
    stop = -(-(rank + 1) * n_items // world)
    return start, min(stop, n_items)


def owner_of(i, n_items, world):
    return i * world // n_items
