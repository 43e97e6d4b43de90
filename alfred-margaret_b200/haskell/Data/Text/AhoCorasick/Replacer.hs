-- | Drop-in for "Data.Text.AhoCorasick.Replacer" (reference: src/Data/Text/AhoCorasick/Replacer.hs:14-27).
-- NOT COMPILED HERE -- see INTEGRATION.md.
module Data.Text.AhoCorasick.Replacer
  ( Replacer, Needle, Replacement, build, compose, mapReplacement, run, runWithLimit
  , setCaseSensitivity, replacerCaseSensitivity
  ) where

import Data.Maybe (fromJust)
import Data.Text.CaseSensitivity (CaseSensitivity (..))
import Data.Text.Utf8 (CodeUnitIndex (..), Text)
import Foreign.ForeignPtr (ForeignPtr, newForeignPtr, withForeignPtr)
import Foreign.Marshal (alloca, withArray)
import Foreign.Ptr (nullPtr)
import Foreign.Storable (peek)
import System.IO.Unsafe (unsafePerformIO)

import qualified Data.Text.AhoCorasick.Automaton as Aho
import qualified Data.Text.Utf8 as Utf8
import Data.Text.AhoCorasick.FFI

type Needle = Text
type Replacement = Text

data Replacer = Replacer                              -- Replacer.hs:78-80 (the Searcher lives in the device handle)
  { replacerCase :: CaseSensitivity
  , replacerPairs :: [(Needle, Replacement)]
  , replacerHandle :: ForeignPtr AmReplacer
  }

-- | `build` (:97-116): pair i gets priority -i; for IgnoreCase the LIBRARY lowers the needles with the
-- host's Char.toLower table and keeps the original byte / code point lengths (:105-113).
build :: CaseSensitivity -> [(Needle, Replacement)] -> Replacer
build cs pairs = unsafePerformIO $
  withArray (map (Aho.toSlice . fst) pairs) $ \ns -> withArray (map (Aho.toSlice . snd) pairs) $ \rs ->
  Aho.withLowerTable IgnoreCase $ \lowerPtr -> alloca $ \out -> do
    rc <- c_am_replacer_build ns rs (fromIntegral (length pairs)) (Aho.caseToC cs) lowerPtr nullPtr out
    if rc /= amOk then Aho.amError "am_replacer_build" else Replacer cs pairs <$> (peek out >>= newForeignPtr c_am_replacer_free_ptr)

compose :: Replacer -> Replacer -> Maybe Replacer                                     -- :120-133
compose a b
  | replacerCase a /= replacerCase b = Nothing
  | otherwise = Just $ build (replacerCase a) (replacerPairs a ++ replacerPairs b)

mapReplacement :: (Replacement -> Replacement) -> Replacer -> Replacer               -- :136-141
mapReplacement f r = build (replacerCase r) [(n, f x) | (n, x) <- replacerPairs r]

replacerCaseSensitivity :: Replacer -> CaseSensitivity
replacerCaseSensitivity = replacerCase

setCaseSensitivity :: CaseSensitivity -> Replacer -> Replacer                         -- :151-153
setCaseSensitivity cs r = build cs (replacerPairs r)

run :: Replacer -> Text -> Text                                                       -- :200-201
run replacer = fromJust . runWithLimit replacer maxBound

-- | `runWithLimit` (:203-242): all passes run on the device; `Nothing` when the length limit is exceeded.
runWithLimit :: Replacer -> CodeUnitIndex -> Text -> Maybe Text
runWithLimit r (CodeUnitIndex maxLength) text = unsafePerformIO $
  withForeignPtr (replacerHandle r) $ \h -> Aho.withSlice text $ \hay ->
  alloca $ \outPtr -> alloca $ \outLen -> alloca $ \exceeded -> do
    let limit = if maxLength == maxBound then maxBound else fromIntegral maxLength
    rc <- c_am_replacer_run h hay limit outPtr outLen exceeded
    if rc /= amOk then Aho.amError "am_replacer_run" else do
      ex <- peek exceeded
      if ex /= 0 then pure Nothing else do
        p <- peek outPtr; n <- peek outLen
        t <- Utf8.fromPtr p (fromIntegral n)      -- copy into a fresh ByteArray#, then
        c_am_free p                               -- release the library's buffer
        pure (Just t)
