{-# LANGUAGE ForeignFunctionInterface #-}
-- | Raw bindings to libam_b200.so (include/am_b200.h).
--
-- NOT COMPILED IN THIS REPOSITORY: GHC is not available in the build image.  This is the shim a maintainer
-- of alfred-margaret would add; it follows the reference's own FFI precedent,
-- benchmark/rust-ffi/app/Main.hs:28-52 (`U8Slice`, pinned arrays, `ccall`).
-- Long GPU calls use `ccall safe` so other Haskell threads keep running.
module Data.Text.AhoCorasick.FFI
  ( U8Slice (..), AmMatch (..), AmLowerPair (..)
  , AmAutomaton, AmReplacer
  , c_am_automaton_build, c_am_automaton_free_ptr
  , c_am_contains_any, c_am_count_matches, c_am_find_all, c_am_contains_all
  , c_am_replacer_build, c_am_replacer_free_ptr, c_am_replacer_run, c_am_free
  , c_am_last_error
  , amOk, amEOverflow
  ) where

import Data.Word (Word8, Word32, Word64)
import Foreign.C.String (CString)
import Foreign.C.Types (CInt (..), CSize (..))
import Foreign.Ptr (FunPtr, Ptr)
import Foreign.Storable (Storable (..))

-- | `am_u8slice`: an unpacked `Text u8data off len`; the array must be pinned
-- (Utf8.isArrayPinned / arrayContents, src/Data/Text/Utf8.hs:325-331).
data U8Slice = U8Slice !(Ptr Word8) !Int !Int

instance Storable U8Slice where
  sizeOf _ = 24
  alignment _ = 8
  peek p = U8Slice <$> peekByteOff p 0 <*> (fromIntegral <$> (peekByteOff p 8 :: IO Word64)) <*> (fromIntegral <$> (peekByteOff p 16 :: IO Word64))
  poke p (U8Slice ptr off len) = pokeByteOff p 0 ptr >> pokeByteOff p 8 (fromIntegral off :: Word64) >> pokeByteOff p 16 (fromIntegral len :: Word64)

-- | `am_match`: (code unit index one past the match, needle index).
data AmMatch = AmMatch !Word64 !Word32

instance Storable AmMatch where
  sizeOf _ = 16
  alignment _ = 8
  peek p = AmMatch <$> peekByteOff p 0 <*> peekByteOff p 8
  poke p (AmMatch e i) = pokeByteOff p 0 e >> pokeByteOff p 8 i >> pokeByteOff p 12 (0 :: Word32)

-- | `am_lower_pair`: one non-identity pair of `Data.Char.toLower` above ASCII.
data AmLowerPair = AmLowerPair !Word32 !Word32

instance Storable AmLowerPair where
  sizeOf _ = 8
  alignment _ = 4
  peek p = AmLowerPair <$> peekByteOff p 0 <*> peekByteOff p 4
  poke p (AmLowerPair a b) = pokeByteOff p 0 a >> pokeByteOff p 4 b

data AmAutomaton
data AmReplacer

amOk, amEOverflow :: CInt
amOk = 0
amEOverflow = 4

-- struct am_lower_table { const am_lower_pair* pairs; size_t n; } and struct am_options are passed by pointer
-- (Ptr ()) from small `alloca`'d buffers in the wrapper modules.
foreign import ccall safe "am_automaton_build"
  c_am_automaton_build :: Ptr U8Slice -> CSize -> CInt -> Ptr () -> Ptr () -> Ptr (Ptr AmAutomaton) -> IO CInt
foreign import ccall "&am_automaton_free"
  c_am_automaton_free_ptr :: FunPtr (Ptr AmAutomaton -> IO ())
-- U8Slice is passed BY VALUE in the C ABI (24 bytes => in memory on SysV x86-64); GHC's FFI cannot pass structs by
-- value, so the shim links a 10-line C file (am_shim.c, see INTEGRATION.md) that takes `const am_u8slice*`.
foreign import ccall safe "am_shim_contains_any"
  c_am_contains_any :: Ptr AmAutomaton -> Ptr U8Slice -> Ptr CInt -> IO CInt
foreign import ccall safe "am_shim_count_matches"
  c_am_count_matches :: Ptr AmAutomaton -> Ptr U8Slice -> Ptr Word64 -> IO CInt
foreign import ccall safe "am_shim_find_all"
  c_am_find_all :: Ptr AmAutomaton -> Ptr U8Slice -> Ptr AmMatch -> CSize -> Ptr Word64 -> IO CInt
foreign import ccall safe "am_shim_contains_all"
  c_am_contains_all :: Ptr AmAutomaton -> Ptr U8Slice -> Ptr CInt -> IO CInt
foreign import ccall safe "am_replacer_build"
  c_am_replacer_build :: Ptr U8Slice -> Ptr U8Slice -> CSize -> CInt -> Ptr () -> Ptr () -> Ptr (Ptr AmReplacer) -> IO CInt
foreign import ccall "&am_replacer_free"
  c_am_replacer_free_ptr :: FunPtr (Ptr AmReplacer -> IO ())
foreign import ccall safe "am_shim_replacer_run"
  c_am_replacer_run :: Ptr AmReplacer -> Ptr U8Slice -> Word64 -> Ptr (Ptr Word8) -> Ptr Word64 -> Ptr CInt -> IO CInt
foreign import ccall unsafe "am_free"
  c_am_free :: Ptr a -> IO ()
foreign import ccall unsafe "am_last_error"
  c_am_last_error :: IO CString
