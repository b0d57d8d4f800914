"""Drop-in mirror of the reference's ``pipeline`` hot-path modules (metrics, enhancement,
dicom_io.normalize_image, schemas)."""
