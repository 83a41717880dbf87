"""csvplus_b200 — B200-native columnar CSV ETL hot path with the csvplus API names.

Product path = csvplus_b200/libcsvplus_b200.so (hand-written sm_100a CUDA kernels behind the C ABI
of include/csvplus_b200.h).  This package is the thin host mirror of the reference's Go API
(csvplus.go) over that ABI; it never computes results on the CPU and fails loudly when the CUDA
library or a GPU is missing.
"""
from .api import (All, Any, Context, CsvPlusError, DataSource, DataSourceError, DeviceBuffer, FromBytes, FromFile,  # noqa: F401
                  FromReadCloser, FromReader, HostBuffer, Index, Like, Not, Predicate, Reader, Row, StopIterationEOF, Table,
                  Take, TakeRows, TakeTable, csv_quote_parity, parse_csv, parse_csv_shard)

__all__ = ["All", "Any", "Context", "CsvPlusError", "DataSource", "DataSourceError", "DeviceBuffer", "FromBytes", "FromFile",
           "FromReadCloser", "FromReader", "HostBuffer", "Index", "Like", "Not", "Predicate", "Reader", "Row",
           "StopIterationEOF", "Table", "Take", "TakeRows", "TakeTable", "csv_quote_parity", "parse_csv", "parse_csv_shard"]
