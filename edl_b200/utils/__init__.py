"""Cluster-state model, coordination primitives and the elastic launcher (reference: python/edl/utils)."""
