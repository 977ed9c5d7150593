from pagraph_b200.partition import dg, hash, utils  # noqa: F401,A004
