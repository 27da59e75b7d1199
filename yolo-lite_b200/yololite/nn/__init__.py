"""Model graph (tasks), backend wrapper (autobackend) and the nn module classes of the YOLO11 path."""
