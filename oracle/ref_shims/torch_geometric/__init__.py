"""Minimal stand-in for torch_geometric (test-only)."""
