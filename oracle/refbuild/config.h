/* Hand-written build configuration for compiling the read-only reference
 * (/root/reference, FFmpeg snapshot, libavcodec 52.78) into oracle/_ref/.
 * It stands in for what the reference's `configure` would generate: pure-C
 * float path (no asm), little-endian x86-64 Linux, only the AAC codec enabled.
 * TEST INFRASTRUCTURE ONLY -- never part of the product. */
#ifndef ORACLE_REFBUILD_CONFIG_H
#define ORACLE_REFBUILD_CONFIG_H
#define FFMPEG_CONFIGURATION "oracle/refbuild (hand-written, --disable-asm equivalent)"
#define FFMPEG_LICENSE "LGPL version 2.1 or later"
#define FFMPEG_DATADIR "/nonexistent"
#define CC_TYPE "gcc"
#define CC_VERSION __VERSION__
#define restrict restrict
#define EXTERN_PREFIX ""
#define EXTERN_ASM
#define SLIBSUF ".so"

#define ARCH_ALPHA 0
#define ARCH_ARM 0
#define ARCH_AVR32 0
#define ARCH_BFIN 0
#define ARCH_MIPS 0
#define ARCH_PPC 0
#define ARCH_SH4 0
#define ARCH_TOMI 0
#define ARCH_X86 0
#define ARCH_X86_32 0
#define ARCH_X86_64 0

#define HAVE_ALTIVEC 0
#define HAVE_MMX 0
#define HAVE_MMI 0
#define HAVE_NEON 0
#define HAVE_VIS 0
#define HAVE_BIGENDIAN 0
#define HAVE_THREADS 0
#define HAVE_6REGS 0
#define HAVE_7REGS 0
#define HAVE_EBX_AVAILABLE 0
#define HAVE_EBP_AVAILABLE 0
#define HAVE_FAST_UNALIGNED 0
#define HAVE_FAST_64BIT 1
#define HAVE_FAST_CMOV 0
#define HAVE_FAST_CLZ 0
#define HAVE_ATTRIBUTE_PACKED 1
#define HAVE_LOCAL_ALIGNED_8 1
#define HAVE_LOCAL_ALIGNED_16 1
#define HAVE_SYMVER_GNU_ASM 0
#define HAVE_SYMVER_ASM_LABEL 0
#define HAVE_MALLOC_H 1
#define HAVE_MEMALIGN 1
#define HAVE_POSIX_MEMALIGN 1
#define HAVE_MKSTEMP 1
#define HAVE_ISATTY 1
#define HAVE_GETHRTIME 0
#define HAVE_EXP2 1
#define HAVE_EXP2F 1
#define HAVE_LLRINT 1
#define HAVE_LLRINTF 1
#define HAVE_LOG2 1
#define HAVE_LOG2F 1
#define HAVE_LRINT 1
#define HAVE_LRINTF 1
#define HAVE_ROUND 1
#define HAVE_ROUNDF 1
#define HAVE_TRUNCF 1

#define CONFIG_HARDCODED_TABLES 0
#define CONFIG_SMALL 0
#define CONFIG_GRAY 0
#define CONFIG_GPL 0
#define CONFIG_MEMALIGN_HACK 0
#define CONFIG_FASTDIV 1
#define CONFIG_MLIB 0
#define CONFIG_MDCT 1
#define CONFIG_LPC 1
#define CONFIG_DWT 0
#define CONFIG_FLOAT 1
#define CONFIG_MPEGAUDIO_HP 0
#define CONFIG_AUDIO_NONSHORT 0
#define CONFIG_ENCODERS 1
#define CONFIG_DECODERS 1
#define CONFIG_AAC_DECODER 1
#define CONFIG_AAC_ENCODER 1
#define CONFIG_AAC_PARSER 1

#define CONFIG_AC3_DECODER 0
#define CONFIG_BINK_DECODER 0
#define CONFIG_CAVS_DECODER 0
#define CONFIG_EATGQ_DECODER 0
#define CONFIG_H261_DECODER 0
#define CONFIG_H261_ENCODER 0
#define CONFIG_H263_DECODER 0
#define CONFIG_H263_ENCODER 0
#define CONFIG_H264_DECODER 0
#define CONFIG_MLP_DECODER 0
#define CONFIG_MSMPEG4_DECODER 0
#define CONFIG_MSMPEG4_ENCODER 0
#define CONFIG_MSMPEG4V1_DECODER 0
#define CONFIG_MSMPEG4V1_ENCODER 0
#define CONFIG_MSMPEG4V2_DECODER 0
#define CONFIG_MSMPEG4V2_ENCODER 0
#define CONFIG_MSMPEG4V3_DECODER 0
#define CONFIG_MSMPEG4V3_ENCODER 0
#define CONFIG_PNG_DECODER 0
#define CONFIG_RV30_DECODER 0
#define CONFIG_RV40_DECODER 0
#define CONFIG_SNOW_ENCODER 0
#define CONFIG_TRUEHD_DECODER 0
#define CONFIG_VC1_DECODER 0
#define CONFIG_VORBIS_DECODER 0
#define CONFIG_VP3_DECODER 0
#define CONFIG_VP5_DECODER 0
#define CONFIG_VP6_DECODER 0
#define CONFIG_VP8_DECODER 0
#define CONFIG_WMV2_DECODER 0
#define CONFIG_WMV2_ENCODER 0
#endif
