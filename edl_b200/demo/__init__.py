"""Demo tooling: JobServer / JobClient (pod rescale schedule driver)."""
