//! felics-gpu -- the public API of visanalexandru/felics (`src/compression.rs`, `src/compression/{traits,format,error}.rs`)
//! on top of the B200 engine's C ABI (include/felics_b200.h).
//!
//! Same names, same signatures, same error behaviour as the reference crate:
//!   * `trait CompressDecompress` (traits.rs:47-65) for `ImageBuffer<Luma<u8|u16>, Vec<_>>` and `ImageBuffer<Rgb<u8|u16>, Vec<_>>`
//!     (compression.rs:250-315, :317-410),
//!   * `compress_image` / `decompress_image` (compression.rs:412-441),
//!   * `read_header` / `write_header` / `Header` / `ColorType` / `PixelDepth` (format.rs:8-84),
//!   * `DecompressionError` (error.rs:5-25).
//! The per-channel loops (compress_channel / decompress_channel, compression.rs:76-248) run as CUDA kernels; there is no
//! CPU fallback: without a usable device every call fails (`io::ErrorKind::Other` / `DecompressionError::IoError`).
//!
//! NOT COMPILED IN THIS REPOSITORY (no Rust toolchain in the build image); see README.md.
pub mod ffi;

use byteorder::{BigEndian, ReadBytesExt, WriteBytesExt};
use image::{DynamicImage, ImageBuffer, Luma, Rgb};
use std::cell::RefCell;
use std::ffi::CStr;
use std::io::{self, Read, Write};

// ---- compression/error.rs:5-25 -------------------------------------------------------------------
#[derive(Debug)]
pub enum DecompressionError {
    IoError(io::Error),
    InvalidValue,
    ValueOverflow,
    InvalidDimensions,
    InvalidColorType,
    InvalidPixelDepth,
    InvalidSignature,
}

impl From<io::Error> for DecompressionError {
    fn from(err: io::Error) -> DecompressionError {
        DecompressionError::IoError(err)
    }
}

// ---- compression/format.rs:8-84 ------------------------------------------------------------------
#[derive(Debug, PartialEq, Eq, Clone, Copy)]
pub enum ColorType {
    Gray = 0,
    Rgb = 1,
}

impl TryFrom<u8> for ColorType {
    type Error = DecompressionError;
    fn try_from(value: u8) -> Result<Self, Self::Error> {
        match value {
            0 => Ok(ColorType::Gray),
            1 => Ok(ColorType::Rgb),
            _ => Err(DecompressionError::InvalidColorType),
        }
    }
}

#[derive(Debug, PartialEq, Eq, Clone, Copy)]
pub enum PixelDepth {
    Eight = 0,
    Sixteen = 1,
}

impl TryFrom<u8> for PixelDepth {
    type Error = DecompressionError;
    fn try_from(value: u8) -> Result<Self, Self::Error> {
        match value {
            0 => Ok(PixelDepth::Eight),
            1 => Ok(PixelDepth::Sixteen),
            _ => Err(DecompressionError::InvalidPixelDepth),
        }
    }
}

pub struct Header {
    pub color_type: ColorType,
    pub pixel_depth: PixelDepth,
    pub width: u32,
    pub height: u32,
}

const MAGIC: [u8; 4] = *b"FLCS";

/// format.rs:51-61
pub fn write_header<W: Write>(header: Header, mut to: W) -> io::Result<()> {
    to.write_all(&MAGIC)?;
    to.write_u8(header.color_type as u8)?;
    to.write_u8(header.pixel_depth as u8)?;
    to.write_u32::<BigEndian>(header.width)?;
    to.write_u32::<BigEndian>(header.height)?;
    Ok(())
}

/// format.rs:63-84 (check order: signature, colour type, pixel depth, then the two u32)
pub fn read_header<R: Read>(mut from: R) -> Result<Header, DecompressionError> {
    let mut magic = [0u8; 4];
    from.read_exact(&mut magic)?;
    if magic != MAGIC {
        return Err(DecompressionError::InvalidSignature);
    }
    let color_type: ColorType = from.read_u8()?.try_into()?;
    let pixel_depth: PixelDepth = from.read_u8()?.try_into()?;
    let width = from.read_u32::<BigEndian>()?;
    let height = from.read_u32::<BigEndian>()?;
    Ok(Header { color_type, pixel_depth, width, height })
}

// ---- compression/traits.rs:7-43 (the constants live in the kernels; only the depth tag is needed on the host) -----
pub trait Intensity: Into<i32> + TryFrom<i32> + Default + Clone + Copy + image::Primitive {
    const PIXEL_DEPTH: PixelDepth;
}
impl Intensity for u8 {
    const PIXEL_DEPTH: PixelDepth = PixelDepth::Eight;
}
impl Intensity for u16 {
    const PIXEL_DEPTH: PixelDepth = PixelDepth::Sixteen;
}

// ---- the engine context: one per thread (calls on one felics_ctx are serialised by the caller, felics_b200.h) ------
struct Ctx(*mut ffi::felics_ctx);

impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { ffi::felics_ctx_destroy(self.0) }
    }
}

thread_local! {
    static CTX: RefCell<Option<Ctx>> = RefCell::new(None);
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::felics_last_error()).to_string_lossy().into_owned() }
}

fn with_ctx<T>(f: impl FnOnce(*mut ffi::felics_ctx) -> T) -> io::Result<T> {
    CTX.with(|slot| {
        let mut slot = slot.borrow_mut();
        if slot.is_none() {
            let mut raw = std::ptr::null_mut();
            let rc = unsafe { ffi::felics_ctx_create(-1, &mut raw) };
            if rc != ffi::FELICS_OK {
                return Err(io::Error::new(io::ErrorKind::Other, format!("felics_b200: {} (no CPU fallback)", last_error())));
            }
            *slot = Some(Ctx(raw));
        }
        Ok(f(slot.as_ref().unwrap().0))
    })
}

fn to_error(rc: i32) -> DecompressionError {
    match rc {
        ffi::FELICS_ERR_IO => DecompressionError::IoError(io::Error::new(io::ErrorKind::UnexpectedEof, "unexpected end of file")),
        ffi::FELICS_ERR_INVALID_VALUE => DecompressionError::InvalidValue,
        ffi::FELICS_ERR_VALUE_OVERFLOW => DecompressionError::ValueOverflow,
        ffi::FELICS_ERR_INVALID_DIMENSIONS => DecompressionError::InvalidDimensions,
        ffi::FELICS_ERR_INVALID_COLOR_TYPE => DecompressionError::InvalidColorType,
        ffi::FELICS_ERR_INVALID_PIXEL_DEPTH => DecompressionError::InvalidPixelDepth,
        ffi::FELICS_ERR_INVALID_SIGNATURE => DecompressionError::InvalidSignature,
        // the reference panics on these streams (parameter_selection.rs:72, rice_coding.rs:50); a library call reports them
        ffi::FELICS_ERR_CORRUPT => DecompressionError::IoError(io::Error::new(io::ErrorKind::InvalidData, "corrupt stream")),
        _ => DecompressionError::IoError(io::Error::new(io::ErrorKind::Other, format!("felics_b200 error {rc}: {}", last_error()))),
    }
}

fn compress_raw<W: Write>(mut to: W, pixels: *const std::ffi::c_void, hdr: ffi::felics_header) -> io::Result<()> {
    // typical streams are far below the input size; FELICS_ERR_BUFFER_TOO_SMALL returns the exact size needed
    let mut cap = unsafe { ffi::felics_pixel_bytes(&hdr) } + ffi::FELICS_HEADER_BYTES + 64;
    loop {
        let mut out = vec![0u8; cap];
        let mut len = 0usize;
        let rc = with_ctx(|ctx| unsafe { ffi::felics_compress(ctx, pixels, &hdr, out.as_mut_ptr(), cap, &mut len) })?;
        match rc {
            ffi::FELICS_OK => return to.write_all(&out[..len]),
            ffi::FELICS_ERR_BUFFER_TOO_SMALL => cap = len + 16,
            _ => return Err(io::Error::new(io::ErrorKind::Other, format!("felics_b200 error {rc}: {}", last_error()))),
        }
    }
}

/// Reads the rest of `from` (the bit stream behind the header) and decodes it into `T` samples.
fn decompress_raw<R: Read, T: Intensity>(mut from: R, header: &Header, channels: usize) -> Result<Vec<T>, DecompressionError> {
    let mut fel = Vec::with_capacity(1 << 16);
    write_header(Header { color_type: header.color_type, pixel_depth: header.pixel_depth, width: header.width, height: header.height }, &mut fel)?;
    from.read_to_end(&mut fel)?;
    let total = (header.width as usize).checked_mul(header.height as usize).ok_or(DecompressionError::InvalidDimensions)?;
    let mut pixels = vec![T::default(); total * channels];
    let mut hdr_out = ffi::felics_header::default();
    let rc = with_ctx(|ctx| unsafe {
        ffi::felics_decompress(ctx, fel.as_ptr(), fel.len(), pixels.as_mut_ptr() as *mut _, pixels.len() * std::mem::size_of::<T>(), &mut hdr_out)
    })?;
    if rc != ffi::FELICS_OK {
        return Err(to_error(rc));
    }
    Ok(pixels)
}

// ---- compression/traits.rs:47-65 ------------------------------------------------------------------
pub trait CompressDecompress {
    fn compress<W>(&self, to: W) -> io::Result<()>
    where
        W: Write;

    fn decompress_with_header<R>(from: R, header: &Header) -> Result<Self, DecompressionError>
    where
        Self: Sized,
        R: Read;

    fn decompress<R>(mut from: R) -> Result<Self, DecompressionError>
    where
        Self: Sized,
        R: Read,
    {
        let header = read_header(&mut from)?;
        Self::decompress_with_header(from, &header)
    }
}

/// compression.rs:250-315
impl<T> CompressDecompress for ImageBuffer<Luma<T>, Vec<T>>
where
    Luma<T>: image::Pixel<Subpixel = T>,
    T: Intensity,
{
    fn compress<W: Write>(&self, to: W) -> io::Result<()> {
        let hdr = ffi::felics_header { color_type: ColorType::Gray as u8, pixel_depth: T::PIXEL_DEPTH as u8, width: self.width(), height: self.height() };
        compress_raw(to, self.as_raw().as_ptr() as *const _, hdr)
    }

    fn decompress_with_header<R: Read>(from: R, header: &Header) -> Result<Self, DecompressionError> {
        if header.color_type != ColorType::Gray {
            return Err(DecompressionError::InvalidColorType);
        }
        if header.pixel_depth != T::PIXEL_DEPTH {
            return Err(DecompressionError::InvalidPixelDepth);
        }
        let pixels = decompress_raw::<R, T>(from, header, 1)?;
        Ok(ImageBuffer::from_raw(header.width, header.height, pixels).unwrap())
    }
}

/// compression.rs:317-410 (the YCoCg-R transform, color_transform.rs:11-26, runs inside the kernels)
impl<T> CompressDecompress for ImageBuffer<Rgb<T>, Vec<T>>
where
    Rgb<T>: image::Pixel<Subpixel = T>,
    T: Intensity,
{
    fn compress<W: Write>(&self, to: W) -> io::Result<()> {
        let hdr = ffi::felics_header { color_type: ColorType::Rgb as u8, pixel_depth: T::PIXEL_DEPTH as u8, width: self.width(), height: self.height() };
        compress_raw(to, self.as_raw().as_ptr() as *const _, hdr)
    }

    fn decompress_with_header<R: Read>(from: R, header: &Header) -> Result<Self, DecompressionError> {
        if header.color_type != ColorType::Rgb {
            return Err(DecompressionError::InvalidColorType);
        }
        if header.pixel_depth != T::PIXEL_DEPTH {
            return Err(DecompressionError::InvalidPixelDepth);
        }
        let pixels = decompress_raw::<R, T>(from, header, 3)?;
        Ok(ImageBuffer::from_raw(header.width, header.height, pixels).unwrap())
    }
}

/// compression.rs:412-418
pub fn compress_image<W, T>(to: W, image: T) -> io::Result<()>
where
    W: Write,
    T: CompressDecompress,
{
    image.compress(to)
}

/// compression.rs:420-441
pub fn decompress_image<R>(mut from: R) -> Result<DynamicImage, DecompressionError>
where
    R: Read,
{
    let header = read_header(&mut from)?;
    let result = match (&header.color_type, &header.pixel_depth) {
        (ColorType::Gray, PixelDepth::Eight) => DynamicImage::ImageLuma8(CompressDecompress::decompress_with_header(from, &header)?),
        (ColorType::Gray, PixelDepth::Sixteen) => DynamicImage::ImageLuma16(CompressDecompress::decompress_with_header(from, &header)?),
        (ColorType::Rgb, PixelDepth::Eight) => DynamicImage::ImageRgb8(CompressDecompress::decompress_with_header(from, &header)?),
        (ColorType::Rgb, PixelDepth::Sixteen) => DynamicImage::ImageRgb16(CompressDecompress::decompress_with_header(from, &header)?),
    };
    Ok(result)
}

#[cfg(test)]
mod test {
    // The reference's own image-level tests (compression.rs:457-558), unchanged in spirit: round trips on odd sizes.
    use super::*;
    use rand::Rng;

    fn round_trip_luma8(width: u32, height: u32) {
        let mut rng = rand::thread_rng();
        let raw: Vec<u8> = (0..width * height).map(|_| rng.gen()).collect();
        let image: ImageBuffer<Luma<u8>, Vec<u8>> = ImageBuffer::from_raw(width, height, raw).unwrap();
        let mut fel = Vec::new();
        image.compress(&mut fel).unwrap();
        let back: ImageBuffer<Luma<u8>, Vec<u8>> = CompressDecompress::decompress(fel.as_slice()).unwrap();
        assert_eq!(image, back);
    }

    #[test]
    fn test_compression_decompression_grayscale() {
        for (w, h) in [(4, 4), (2, 1), (1, 2), (1, 1), (1447, 8), (1, 100), (0, 5)] {
            round_trip_luma8(w, h);
        }
    }
}
