"""Drop-in for the reference package models/ops (see models/__init__.py)."""
