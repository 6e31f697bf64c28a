"""Multi-GPU: one process per GPU, tables sharded by feature field (placeholder, filled in below)."""


def attach(model, device):
    raise NotImplementedError("multi-GPU field sharding is not wired yet")
