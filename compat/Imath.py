"""Inert placeholder."""
