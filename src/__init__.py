"""Drop-in for the reference's ``src`` package: same module names and entry points
(``python -m src.mnist``, ``python -m src.shakespeare``), B200 kernels underneath."""
try:  # the reference calls load_dotenv() here (src/__init__.py:1-2); optional in this build
    from dotenv import load_dotenv

    load_dotenv()
except Exception:  # pragma: no cover - dotenv absent or unreadable .env
    pass
