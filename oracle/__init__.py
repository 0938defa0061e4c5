"""ORACLE — test infrastructure (CPU fp32 restatement of the hot path). PARITY UNPINNED; see graphs.py."""
