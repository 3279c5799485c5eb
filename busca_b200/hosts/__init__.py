"""Host-tracker drivers: from-scratch callers that use BUSCA exactly as the reference's adapters do (same call pattern,
same arguments), for tests, benchmarks and as usage examples.  The arithmetic of the hot path stays in libbusca_b200.so."""
