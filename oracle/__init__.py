"""Test infrastructure: ctypes access to the CPU restatement (oracle/pt_oracle.c) and to the reference
binaries built by oracle/ref/Makefile.  Product code (tungsten_b200/) must never import this package."""
