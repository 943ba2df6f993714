"""Drop-in shim for the reference's `robot` package (robot/robot.py constants and exception type)."""
