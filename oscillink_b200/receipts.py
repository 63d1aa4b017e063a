"""Host-side receipt verification (HMAC-SHA256), mirroring oscillink/core/receipts.py:86-179.

Signing and verification stay on the host by design (BASELINE.json north_star); nothing here
touches the GPU.
"""
from __future__ import annotations

import hashlib
import hmac
import json

__all__ = ["verify_receipt", "verify_receipt_mode"]


def _as_bytes(secret: bytes | str) -> bytes:
    return secret.encode("utf-8") if isinstance(secret, str) else secret


def _digest(payload: dict, secret: bytes) -> str:
    return hmac.new(secret, json.dumps(payload, sort_keys=True).encode("utf-8"), hashlib.sha256).hexdigest()


def _signature_block(receipt: dict):
    block = receipt.get("meta", {}).get("signature")
    if not block or block.get("algorithm") != "HMAC-SHA256":
        return None
    if block.get("payload") is None or block.get("signature") is None:
        return None
    return block


def verify_receipt(receipt: dict, secret: bytes | str) -> bool:
    """True iff receipt['meta']['signature'] carries a valid HMAC of its payload.  Never raises."""
    try:
        block = _signature_block(receipt)
        if block is None:
            return False
        return hmac.compare_digest(_digest(block["payload"], _as_bytes(secret)), str(block["signature"]))
    except Exception:
        return False


def verify_receipt_mode(
    receipt: dict,
    secret: bytes | str,
    require_mode: str | None = None,
    minimal_subset: bool = False,
    required_sig_v: int | None = None,
) -> tuple[bool, dict | None]:
    """Verification with mode / version requirements and the minimal-subset bridge
    (receipts.py:113-179): returns (ok, payload_or_None) and never raises."""
    try:
        block = _signature_block(receipt)
        if block is None:
            return False, None
        payload, claimed = block["payload"], str(block["signature"])
        mode = payload.get("mode")
        if require_mode and mode != require_mode:
            return False, None
        if required_sig_v is not None and payload.get("sig_v") != required_sig_v:
            return False, None
        key = _as_bytes(secret)
        if hmac.compare_digest(_digest(payload, key), claimed):
            return True, payload
        if minimal_subset and mode == "extended":
            core = {"sig_v": payload.get("sig_v"), "mode": "minimal",
                    "state_sig": payload.get("state_sig"), "deltaH_total": payload.get("deltaH_total")}
            if hmac.compare_digest(_digest(core, key), claimed) and require_mode in (None, "minimal"):
                return True, core
        return False, None
    except Exception:
        return False, None
