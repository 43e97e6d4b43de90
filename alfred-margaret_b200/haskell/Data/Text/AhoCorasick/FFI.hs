{-# LANGUAGE BangPatterns #-}
{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE MagicHash #-}
{-# LANGUAGE UnboxedTuples #-}
-- | Raw bindings to libam_b200.so (include/am_b200.h, ABI version 2) and the marshalling helpers every wrapper
-- module uses.
--
-- NOT COMPILED IN THIS REPOSITORY: GHC is not available in the build image.  This is the shim a maintainer
-- of alfred-margaret would add; it follows the reference's own FFI precedent,
-- benchmark/rust-ffi/app/Main.hs:28-52: `Ptr U8Slice` arguments over pinned arrays, plain integers back.
-- Every struct of the header travels by pointer, so no C glue is needed (GHC's FFI cannot pass structs by value).
-- Long GPU calls are `ccall safe` so other Haskell threads keep running; the error text is thread-local in the
-- library, so a failed call and the read of its message are bound to one OS thread (`amCall`).
module Data.Text.AhoCorasick.FFI
  ( U8Slice (..), AmMatch (..), AmLowerPair (..)
  , AmAutomaton, AmReplacer
  , c_am_automaton_build, c_am_automaton_free_ptr, c_am_automaton_prepare
  , c_am_contains_any, c_am_count_matches, c_am_find_all, c_am_contains_all
  , c_am_replacer_build, c_am_replacer_build_stored, c_am_replacer_free_ptr, c_am_replacer_run, c_am_free
  , amOk, amEOverflow
  , caseToC, toSlice, withSlice, withSlices, withLowerTable, amCall, pinText
  ) where

import Control.Concurrent (runInBoundThread)
import Control.Monad (when)
import Data.Text.CaseSensitivity (CaseSensitivity (..))
import Data.Text.Utf8 (Text (..))
import Data.Word (Word8, Word32, Word64)
import Foreign.C.String (peekCString)
import Foreign.C.Types (CChar, CInt (..), CSize (..))
import Foreign.Marshal (alloca, allocaBytes, withArrayLen)
import Foreign.Ptr (FunPtr, Ptr, castPtr, nullPtr)
import Foreign.Storable (Storable (..))
import System.IO.Unsafe (unsafePerformIO)
import GHC.Exts (touch#)
import GHC.IO (IO (..))

import qualified Data.Char as Char
import qualified Data.Text.Array as TextArray
import qualified Data.Text.Utf8 as Utf8

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

-- | `AM_CASE_SENSITIVE = 0`, `AM_IGNORE_CASE = 1` (Data.Text.CaseSensitivity, :14-16).
caseToC :: CaseSensitivity -> CInt
caseToC CaseSensitive = 0
caseToC IgnoreCase = 1

-- struct am_lower_table { const am_lower_pair* pairs; size_t n; } and struct am_options travel by pointer (Ptr ()).
foreign import ccall safe "am_automaton_build"
  c_am_automaton_build :: Ptr U8Slice -> CSize -> Ptr () -> Ptr () -> Ptr (Ptr AmAutomaton) -> IO CInt
foreign import ccall "&am_automaton_free"
  c_am_automaton_free_ptr :: FunPtr (Ptr AmAutomaton -> IO ())
foreign import ccall safe "am_automaton_prepare"
  c_am_automaton_prepare :: Ptr AmAutomaton -> CInt -> IO CInt
foreign import ccall safe "am_contains_any"
  c_am_contains_any :: Ptr AmAutomaton -> CInt -> Ptr U8Slice -> Ptr CInt -> IO CInt
foreign import ccall safe "am_count_matches"
  c_am_count_matches :: Ptr AmAutomaton -> CInt -> Ptr U8Slice -> Ptr Word64 -> IO CInt
foreign import ccall safe "am_find_all"
  c_am_find_all :: Ptr AmAutomaton -> CInt -> Ptr U8Slice -> Ptr AmMatch -> CSize -> Ptr Word64 -> IO CInt
foreign import ccall safe "am_contains_all"
  c_am_contains_all :: Ptr AmAutomaton -> CInt -> Ptr U8Slice -> Ptr CInt -> IO CInt
foreign import ccall safe "am_replacer_build"
  c_am_replacer_build :: Ptr U8Slice -> Ptr U8Slice -> CSize -> CInt -> Ptr () -> Ptr () -> Ptr (Ptr AmReplacer) -> IO CInt
foreign import ccall safe "am_replacer_build_stored"
  c_am_replacer_build_stored :: Ptr U8Slice -> Ptr Word32 -> Ptr Word32 -> Ptr U8Slice -> CSize -> CInt -> Ptr () -> Ptr () -> Ptr (Ptr AmReplacer) -> IO CInt
foreign import ccall "&am_replacer_free"
  c_am_replacer_free_ptr :: FunPtr (Ptr AmReplacer -> IO ())
foreign import ccall safe "am_replacer_run"
  c_am_replacer_run :: Ptr AmReplacer -> CInt -> Ptr U8Slice -> Word64 -> Ptr (Ptr Word8) -> Ptr Word64 -> Ptr CInt -> IO CInt
foreign import ccall unsafe "am_free"
  c_am_free :: Ptr a -> IO ()
foreign import ccall unsafe "am_last_error_copy"
  c_am_last_error_copy :: Ptr CChar -> CSize -> IO CSize

-- | Run an ABI call and turn a non-zero status (other than the ones in `allowed`) into `error` with the library's
-- message.  The reference has no error channel on this path (its functions are total on valid input), so a failure
-- here -- no device, out of memory -- is exceptional.  Call and message read share one OS thread.
amCall :: String -> [CInt] -> IO CInt -> IO CInt
amCall what allowed act = runInBoundThread $ do
  rc <- act
  when (rc /= amOk && rc `notElem` allowed) $ do
    msg <- allocaBytes 512 $ \buf -> c_am_last_error_copy buf 512 >> peekCString buf
    error (what ++ ": libam_b200 status " ++ show rc ++ ": " ++ msg)
  pure rc

-- | A Text whose array is pinned, so that its address may cross the FFI (Utf8.isArrayPinned, Utf8.hs:325-328).  GHC
-- pins large arrays (>= ~3 KiB) itself; a small unpinned one is copied once into a pinned array, as the reference's
-- benchmark does with `compact` (benchmark/rust-ffi/app/Main.hs:75).
pinText :: Text -> Text
pinText t@(Text u8data off len)
  | Utf8.isArrayPinned u8data = t
  | otherwise = Text (TextArray.run (do { dst <- TextArray.newPinned len; TextArray.copyI len dst 0 u8data off; pure dst })) 0 len

-- | The array's address has crossed the FFI: the garbage collector must not free it before the call returns.
touchArray :: TextArray.Array -> IO ()
touchArray (TextArray.ByteArray ba#) = IO (\s -> case touch# ba# s of s' -> (# s', () #))

-- | `U8Slice` of a PINNED Text (the reference's `fromText`, benchmark/rust-ffi/app/Main.hs:49-52).
toSlice :: Text -> U8Slice
toSlice (Text u8data off len)
  | Utf8.isArrayPinned u8data = U8Slice (Utf8.arrayContents u8data) off len
  | otherwise                 = error "ByteArray is not pinned"

-- | Pin if needed, write the slice to a temporary and keep the array alive across the call.
withSlice :: Text -> (Ptr U8Slice -> IO a) -> IO a
withSlice text act =
  let !pinned@(Text arr _ _) = pinText text
  in alloca $ \p -> do
       poke p (toSlice pinned)
       r <- act p
       touchArray arr           -- keeps the ByteArray# alive until the call has returned
       pure r

-- | An array of slices (needles, replacements).
withSlices :: [Text] -> (Ptr U8Slice -> CSize -> IO a) -> IO a
withSlices texts act =
  let pinned = map pinText texts
  in withArrayLen (map toSlice pinned) $ \n p -> do
       r <- act p (fromIntegral n)
       mapM_ (\(Text arr _ _) -> touchArray arr) pinned
       pure r

-- | The host's `Data.Char.toLower` above ASCII as data (Utf8.lowerCodePoint, Utf8.hs:145-151, is `Char.toLower` there):
-- the table depends on the GHC that compiles the host (Unicode version of `base`), so it is enumerated here, once.
lowerPairs :: [AmLowerPair]
lowerPairs = [ AmLowerPair (fromIntegral (Char.ord c)) (fromIntegral (Char.ord l))
             | c <- ['\x80' .. maxBound], let l = Char.toLower c, l /= c ]
{-# NOINLINE lowerPairs #-}

-- | Pass `am_lower_table { pairs, n }` by pointer.
withLowerTable :: (Ptr () -> IO a) -> IO a
withLowerTable act =
  withArrayLen lowerPairs $ \n pairs ->
  allocaBytes 16 $ \tbl -> do
    pokeByteOff tbl 0 pairs
    pokeByteOff tbl 8 (fromIntegral n :: Word64)
    act (castPtr tbl)
