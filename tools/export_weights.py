"""Export the packed weights of the three engines as ONE relocatable blob for non-Python hosts.

File layout (little endian):
  "VBW1" | int32 n_structs | int64 data_bytes
  per struct: char name[16] | int64 struct_bytes | int64 n_reloc | struct bytes | int64 reloc_offsets[n_reloc]
  data section (every tensor 256-byte aligned)
Every pointer field of a descriptor (vb_dit_desc, vb_dit_prep_desc, vb_vae_desc, vb_voc_desc of include/versband.h)
holds `1 + offset into the data section` (0 = NULL); `reloc_offsets` lists the byte positions of the non-NULL pointer
fields inside the struct. A host loads the data section to the device and adds `device_base - 1` at every relocation
(examples/c_host/run_clip.c).
"""
from __future__ import annotations

import ctypes as C
import struct
from typing import Dict, List, Tuple

import torch


def _pointer_fields_of(ctype, base: int = 0) -> List[int]:
    """Byte offsets of every pointer-typed scalar inside a ctypes type (Structure / Array, recursive)."""
    if ctype is C.c_void_p or (isinstance(ctype, type) and issubclass(ctype, C._Pointer)):
        return [base]
    if isinstance(ctype, type) and issubclass(ctype, C.Array):
        size = C.sizeof(ctype._type_)
        return [o for i in range(ctype._length_) for o in _pointer_fields_of(ctype._type_, base + i * size)]
    if isinstance(ctype, type) and issubclass(ctype, C.Structure):
        return [o for name, ft in ctype._fields_ for o in _pointer_fields_of(ft, base + getattr(ctype, name).offset)]
    return []


def _tensors(obj) -> List[torch.Tensor]:
    if isinstance(obj, torch.Tensor):
        return [obj]
    if isinstance(obj, dict):
        return [t for v in obj.values() for t in _tensors(v)]
    if isinstance(obj, (list, tuple)):
        return [t for v in obj for t in _tensors(v)]
    return []


def export(path: str, structs: List[Tuple[str, C.Structure]], weight_holders) -> Dict[str, int]:
    """structs: [(name, ctypes descriptor)], weight_holders: objects (dicts / lists) holding every tensor the
    descriptors point to."""
    by_ptr = {t.data_ptr(): t for t in _tensors(weight_holders)}
    data = bytearray()
    placed: Dict[int, int] = {}
    blobs = []
    for name, st in structs:
        raw = bytearray(bytes(st))
        relocs = []
        for off in _pointer_fields_of(type(st)):
            ptr = struct.unpack_from("<Q", raw, off)[0]
            if ptr == 0:
                continue
            if ptr not in placed:
                t = by_ptr.get(ptr)
                if t is None:
                    raise KeyError(f"{name}: pointer field at byte {off} does not address the start of a known weight tensor")
                pad = (-len(data)) % 256
                data.extend(b"\0" * pad)
                placed[ptr] = len(data)
                data.extend(t.contiguous().view(torch.uint8).cpu().numpy().tobytes())
            struct.pack_into("<Q", raw, off, placed[ptr] + 1)
            relocs.append(off)
        blobs.append((name, raw, relocs))
    with open(path, "wb") as f:
        f.write(b"VBW1")
        f.write(struct.pack("<iq", len(blobs), len(data)))
        for name, raw, relocs in blobs:
            f.write(name.encode().ljust(16, b"\0"))
            f.write(struct.pack("<qq", len(raw), len(relocs)))
            f.write(raw)
            f.write(struct.pack(f"<{len(relocs)}q", *relocs))
        f.write(data)
    return {"structs": len(blobs), "tensors": len(placed), "data_bytes": len(data)}


def export_pipeline(path: str, pipe) -> Dict[str, int]:
    """All four descriptors of an AccompBandPipeline."""
    dit, vae, voc = pipe.dit, pipe.vae, pipe.voc
    return export(path, [("dit", dit._desc()), ("prep", dit._prep_desc()), ("vae", vae._desc()), ("voc", voc._desc())],
                  [dit.w, vae.w, voc.w])
