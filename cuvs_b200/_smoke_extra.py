"""Additional per-index smoke checks appended as index types land (called by __graft_entry__.smoke)."""


def run():
    pass
