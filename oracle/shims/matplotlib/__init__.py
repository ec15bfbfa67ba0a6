"""import-only stub (reference imports matplotlib.pyplot at module import)."""
