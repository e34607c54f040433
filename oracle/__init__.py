"""ORACLE — test infrastructure only (see oracle/ba_math.hpp). PARITY UNPINNED."""
