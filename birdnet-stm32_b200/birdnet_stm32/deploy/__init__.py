"""Host side of the board test that concerns the hot path: reading the firmware's UART report and comparing it with the
B200 engine on the same files (everything else of `deploy/` -- stedgeai, flashing, project patching -- is out of scope)."""
