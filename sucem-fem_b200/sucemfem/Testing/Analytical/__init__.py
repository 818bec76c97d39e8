"""Closed-form solutions the examples compare against (mirror of /root/reference/sucemfem/Testing/Analytical/)."""
