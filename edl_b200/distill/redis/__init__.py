"""Redis-flavoured teacher registry + balance server (reference: python/edl/distill/redis/*)."""
