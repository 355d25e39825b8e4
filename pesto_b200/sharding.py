"""Structure-level sharding across GPUs: structures are independent, so ranks get disjoint sets of
structures and never communicate on the data path.  Longest-processing-time greedy bin packing by cost."""
import heapq


def lpt_partition(costs, n_ranks):
    """Return n_ranks lists of structure indices with near-equal total cost (cost ~ atoms * sum(nn))."""
    order = sorted(range(len(costs)), key=lambda i: (-int(costs[i]), i))
    heap = [(0, r) for r in range(n_ranks)]
    shards = [[] for _ in range(n_ranks)]
    for i in order:
        load, r = heapq.heappop(heap)
        shards[r].append(i)
        heapq.heappush(heap, (load + int(costs[i]), r))
    return shards


def rank_shard(costs, rank, world_size):
    return lpt_partition(costs, world_size)[rank]
