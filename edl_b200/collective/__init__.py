"""Elastic collective training: launcher CLI, elastic data reader, user state serialisation."""
