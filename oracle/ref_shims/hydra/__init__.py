"""Minimal stand-in for hydra (test-only)."""
