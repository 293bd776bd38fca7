//! Golden vectors from the reference's own arithmetic.
//!
//!   cargo run --release -- <inputs dir> <repo>/tests/golden/golden.json <repo>/tests/golden/ref_golden.json
//!
//! `<inputs dir>` is what `python oracle/pin/dump_inputs.py <inputs dir>` wrote: one raw u8 file per golden case
//! (`<name>.raw`, tightly packed rows, interleaved channels) and `manifest.json` with each case's dims, channels and
//! request parameters.  For every case this program performs, call for call, the pixel section of
//! `State::process_image` (reference src/handler.rs:224-255) or of the per-frame closure of `process_gif`
//! (:329-355) on a `DynamicImage` of the same variant the decoder would have produced, and records the SHA-256 of
//! the resulting raw pixels in golden.json's schema.  `tests/test_oracle.py::test_oracle_matches_reference_golden`
//! compares the C oracle with that file whenever it exists: that is what turns "parity unpinned" into "pinned".
use image::imageops::FilterType;
use image::{DynamicImage, GrayAlphaImage, GrayImage, ImageBuffer, Luma, LumaA, Rgb, Rgba, RgbImage, RgbaImage};
use serde::{Deserialize, Serialize};
use sha2::{Digest, Sha256};
use std::{env, fs, path::Path};

#[derive(Deserialize)]
struct Params {
    w: Option<u32>,
    h: Option<u32>,
    rgb: Option<[u8; 3]>,
    #[serde(default)]
    crop: bool,
    #[serde(default)]
    blur: f32, // already through Query::blur(): 0 or clamp(.,10,20)  (src/query.rs:59-62)
    #[serde(default)]
    grayscale: bool,
    #[serde(default)]
    inverse: bool,
    #[serde(default)]
    gif: bool,
    #[serde(default)]
    orientation: u8, // EXIF value; 0 / 1 = none
    #[serde(default)]
    to_rgb8: bool,
    #[serde(default)]
    to_rgba8: bool, // the WebP branch: img.into_rgba8(), src/handler.rs:287
}

#[derive(Deserialize)]
struct CaseIn {
    name: String,
    width: u32,
    height: u32,
    channels: u32,
    #[serde(default)]
    sample: String, // "u8" (default), "u16", "f32": the subpixel type of the DynamicImage variant (golden_deep.json)
    params: Params,
}

#[derive(Serialize)]
struct CaseOut {
    name: String,
    out_h: u32,
    out_w: u32,
    out_c: u32,
    out_dtype: &'static str, // "uint8" | "uint16" | "float32": subpixel type of the variant the stage ended in
    sha256: String,
}

#[derive(Serialize)]
struct Golden {
    oracle: &'static str,
    pinned_to_reference: bool,
    cases: Vec<CaseOut>,
}

fn load(dir: &Path, c: &CaseIn) -> DynamicImage {
    let raw = fs::read(dir.join(format!("{}.raw", c.name))).expect("raw input");
    let (w, h) = (c.width, c.height);
    if c.sample == "u16" {
        // ImageLuma16 .. ImageRgba16: what a 16-bit PNG / TIFF decodes to (src/handler.rs:219)
        assert_eq!(raw.len() as u32, w * h * c.channels * 2);
        let v: Vec<u16> = raw.chunks_exact(2).map(|b| u16::from_le_bytes([b[0], b[1]])).collect();
        return match c.channels {
            1 => DynamicImage::ImageLuma16(ImageBuffer::<Luma<u16>, _>::from_raw(w, h, v).unwrap()),
            2 => DynamicImage::ImageLumaA16(ImageBuffer::<LumaA<u16>, _>::from_raw(w, h, v).unwrap()),
            3 => DynamicImage::ImageRgb16(ImageBuffer::<Rgb<u16>, _>::from_raw(w, h, v).unwrap()),
            4 => DynamicImage::ImageRgba16(ImageBuffer::<Rgba<u16>, _>::from_raw(w, h, v).unwrap()),
            n => panic!("channels {n}"),
        };
    }
    if c.sample == "f32" {
        // ImageRgb32F / ImageRgba32F: HDR / EXR
        assert_eq!(raw.len() as u32, w * h * c.channels * 4);
        let v: Vec<f32> = raw.chunks_exact(4).map(|b| f32::from_le_bytes([b[0], b[1], b[2], b[3]])).collect();
        return match c.channels {
            3 => DynamicImage::ImageRgb32F(ImageBuffer::<Rgb<f32>, _>::from_raw(w, h, v).unwrap()),
            4 => DynamicImage::ImageRgba32F(ImageBuffer::<Rgba<f32>, _>::from_raw(w, h, v).unwrap()),
            n => panic!("channels {n}"),
        };
    }
    assert_eq!(raw.len() as u32, c.width * c.height * c.channels);
    match c.channels {
        1 => DynamicImage::ImageLuma8(GrayImage::from_raw(c.width, c.height, raw).unwrap()),
        2 => DynamicImage::ImageLumaA8(GrayAlphaImage::from_raw(c.width, c.height, raw).unwrap()),
        3 => DynamicImage::ImageRgb8(RgbImage::from_raw(c.width, c.height, raw).unwrap()),
        4 => DynamicImage::ImageRgba8(RgbaImage::from_raw(c.width, c.height, raw).unwrap()),
        n => panic!("channels {n}"),
    }
}

/// reference src/handler.rs:224-255 (still) and :329-355 (GIF frame), verbatim in structure.
fn stage(mut img: DynamicImage, p: &Params) -> DynamicImage {
    if !p.gif && p.orientation >= 2 {
        // src/handler.rs:206,221-223
        let o = image::metadata::Orientation::from_exif(p.orientation).expect("exif orientation");
        img.apply_orientation(o);
    }
    if p.grayscale {
        img = img.grayscale();
    } else if p.inverse {
        img.invert();
    }
    let filter = if p.gif { FilterType::Nearest } else { FilterType::Lanczos3 };
    if let (Some(w), Some(h)) = (p.w, p.h) {
        if w != img.width() || h != img.height() {
            img = if p.crop { img.resize_to_fill(w, h, filter) } else { img.resize(w, h, filter) };
        }
        if w > img.width() || h > img.height() {
            let [r, g, b] = p.rgb.unwrap_or([32, 32, 32]); // Query::fill_color() default, src/query.rs:35-49
            let mut bg = ImageBuffer::from_pixel(w, h, Rgba([r, g, b, 255]));
            let x = i64::from(w.abs_diff(img.width()) / 2);
            let y = i64::from(h.abs_diff(img.height()) / 2);
            image::imageops::overlay(&mut bg, &img, x, y);
            img = DynamicImage::ImageRgba8(bg);
        }
    }
    if !p.gif && p.blur > 0.0 {
        img = img.blur(p.blur);
    }
    if p.gif {
        img = DynamicImage::ImageRgba8(img.to_rgba8()); // src/handler.rs:355
    }
    if p.to_rgb8 {
        img = DynamicImage::ImageRgb8(img.to_rgb8()); // the JPEG branch, src/handler.rs:274-278
    }
    if p.to_rgba8 && !p.gif {
        img = DynamicImage::ImageRgba8(img.into_rgba8()); // the WebP branch, src/handler.rs:287
    }
    img
}

fn main() {
    let a: Vec<String> = env::args().collect();
    assert!(a.len() == 4, "usage: fanlin-oracle-pin <inputs dir> <golden.json> <ref_golden.json>");
    let dir = Path::new(&a[1]);
    let manifest: Vec<CaseIn> = serde_json::from_slice(&fs::read(dir.join("manifest.json")).unwrap()).unwrap();
    let mut cases = Vec::new();
    for c in &manifest {
        let out = stage(load(dir, c), &c.params);
        let ch = u32::from(out.color().channel_count());
        let out_dtype = match out.color().bytes_per_pixel() / out.color().channel_count() {
            1 => "uint8",
            2 => "uint16",
            _ => "float32",
        };
        let mut h = Sha256::new();
        h.update(out.as_bytes());
        cases.push(CaseOut {
            name: c.name.clone(),
            out_h: out.height(),
            out_w: out.width(),
            out_c: ch,
            out_dtype,
            sha256: h.finalize().iter().map(|b| format!("{b:02x}")).collect(),
        });
    }
    let g = Golden { oracle: "image = 0.25.6 through the call sequence of fanlin-rs src/handler.rs:224-255,329-355", pinned_to_reference: true, cases };
    fs::write(&a[3], serde_json::to_string_pretty(&g).unwrap()).unwrap();
    let _ = &a[2]; // golden.json is only named so that the two files sit next to each other in the command line
    eprintln!("wrote {} cases to {}", manifest.len(), a[3]);
}
