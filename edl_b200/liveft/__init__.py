"""Live fault tolerance: np-matching rendezvous (reference: python/edl/liveft)."""
