"""TIFF comparison helpers (reporting only; reference src/flac_raster/compare.py)."""
from __future__ import annotations

from pathlib import Path

import numpy as np

from .tiffio import read_geotiff


def compare_tiffs(file1_path: Path, file2_path: Path, show_bands: bool = True) -> dict:
    """Same result keys as the reference (compare.py:17-82)."""
    file1_path, file2_path = Path(file1_path), Path(file2_path)
    r1, r2 = read_geotiff(file1_path), read_geotiff(file2_path)
    d1, d2 = r1.data, r2.data
    results = {
        "file1": file1_path.name,
        "file2": file2_path.name,
        "shape_match": d1.shape == d2.shape,
        "dtype_match": d1.dtype == d2.dtype,
        "shape1": d1.shape,
        "shape2": d2.shape,
        "dtype1": str(d1.dtype),
        "dtype2": str(d2.dtype),
    }
    if results["shape_match"]:
        diff = d1.astype(np.float64) - d2.astype(np.float64)
        results.update({
            "arrays_equal": bool(np.array_equal(d1, d2)),
            "max_diff": float(np.nanmax(np.abs(diff))) if diff.size else 0.0,
            "mean_diff": float(np.nanmean(np.abs(diff))) if diff.size else 0.0,
            "different_pixels": int(np.count_nonzero(diff)),
            "total_pixels": int(diff.size),
        })
        if show_bands:
            results["bands"] = [{"band": b + 1, "equal": bool(np.array_equal(d1[b], d2[b])),
                                 "max_diff": float(np.nanmax(np.abs(diff[b])))} for b in range(d1.shape[0])]
    results["transform_match"] = r1.transform == r2.transform
    results["crs_match"] = r1.crs == r2.crs
    return results


def display_comparison_table(results: dict):
    for k, v in results.items():
        if k != "bands":
            print(f"{k:>18}: {v}")
    for b in results.get("bands", []):
        print(f"{'band ' + str(b['band']):>18}: equal={b['equal']} max_diff={b['max_diff']}")
