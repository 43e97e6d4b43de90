-- | Drop-in for "Data.Text.AhoCorasick.Replacer" (reference: src/Data/Text/AhoCorasick/Replacer.hs:14-27).
-- NOT COMPILED HERE -- see INTEGRATION.md.
module Data.Text.AhoCorasick.Replacer
  ( Replacer, Needle, Replacement, Payload (..), build, compose, mapReplacement, run, runWithLimit
  , setCaseSensitivity, replacerCaseSensitivity
  ) where

import Control.Monad.ST (stToIO)
import Data.Maybe (fromJust)
import Data.Text.CaseSensitivity (CaseSensitivity (..))
import Data.Text.Utf8 (CodeUnitIndex (..), Text (..))
import Foreign.ForeignPtr (ForeignPtr, newForeignPtr, withForeignPtr)
import Foreign.Marshal (alloca, withArray)
import Foreign.Ptr (nullPtr)
import Foreign.Storable (peek)
import System.IO.Unsafe (unsafePerformIO)

import qualified Data.Text as Text
import qualified Data.Text.Array as TextArray
import qualified Data.Text.Utf8 as Utf8
import Data.Text.AhoCorasick.FFI

type Needle = Text
type Replacement = Text

data Payload = Payload                                -- Replacer.hs:59-69
  { needlePriority :: !Int
  , needleLengthBytes :: !CodeUnitIndex
  , needleLengthCodePoints :: !Int
  , needleReplacement :: !Replacement
  }

-- | The reference's `Replacer { replacerSearcher :: Searcher Payload }` (:78-80): the case flag, the needles as the
-- searcher stores them (lowered iff built with IgnoreCase, :105-107) with their payloads -- and the device handle
-- that was built from exactly those.
data Replacer = Replacer
  { replacerCase :: CaseSensitivity
  , replacerStored :: [(Needle, Payload)]
  , replacerHandle :: ForeignPtr AmReplacer
  }

-- | `build` (:97-116): pair i gets priority -i; the needles of an IgnoreCase replacer are lowered (`Utf8.lowerUtf8`, :107)
-- and the payload keeps the lengths of the ORIGINAL needle (:111-113).
build :: CaseSensitivity -> [(Needle, Replacement)] -> Replacer
build cs pairs = fromStored cs (zipWith mapNeedle [0 ..] pairs)
  where
    mapNeedle i (needle, replacement) =
      ( case cs of { CaseSensitive -> needle; IgnoreCase -> Utf8.lowerUtf8 needle }
      , Payload (negate i) (Utf8.lengthUtf8 needle) (Text.length needle) replacement )

-- | `Searcher.buildWithValues` over stored pairs: nothing is lowered, the payload lengths are taken as they are
-- (am_replacer_build_stored).  Priorities are renumbered 0, -1, ... in list order, as `build` and `compose` do.
fromStored :: CaseSensitivity -> [(Needle, Payload)] -> Replacer
fromStored cs stored0 = unsafePerformIO $
  withSlices (map fst stored) $ \ns n -> withSlices (map (needleReplacement . snd) stored) $ \rs _ ->
  withArray (map (fromIntegral . codeUnitIndex . needleLengthBytes . snd) stored) $ \lenBytes ->
  withArray (map (fromIntegral . needleLengthCodePoints . snd) stored) $ \lenCps ->
  withLowerTable $ \lowerPtr -> alloca $ \out -> do
    _ <- amCall "am_replacer_build_stored" [] $ c_am_replacer_build_stored ns lenBytes lenCps rs n (caseToC cs) lowerPtr nullPtr out
    Replacer cs stored <$> (peek out >>= newForeignPtr c_am_replacer_free_ptr)
  where stored = zipWith (\i (nd, p) -> (nd, p { needlePriority = negate i })) [0 ..] stored0
{-# NOINLINE fromStored #-}

compose :: Replacer -> Replacer -> Maybe Replacer                                     -- :120-133: the STORED needles, renumbered
compose a b
  | replacerCase a /= replacerCase b = Nothing
  | otherwise = Just $ fromStored (replacerCase a) (replacerStored a ++ replacerStored b)

mapReplacement :: (Replacement -> Replacement) -> Replacer -> Replacer               -- :136-141: "doesn't modify the needles"
mapReplacement f r = fromStored (replacerCase r) [(n, p { needleReplacement = f (needleReplacement p) }) | (n, p) <- replacerStored r]

replacerCaseSensitivity :: Replacer -> CaseSensitivity
replacerCaseSensitivity = replacerCase

-- | :151-153: "Does not change the capitilization of the needles": only the flag changes, the handle is shared and
-- the run passes the flag.
setCaseSensitivity :: CaseSensitivity -> Replacer -> Replacer
setCaseSensitivity cs r = r { replacerCase = cs }

run :: Replacer -> Text -> Text                                                       -- :200-201
run replacer = fromJust . runWithLimit replacer maxBound

-- | `runWithLimit` (:203-242): all passes run on the device; `Nothing` when the length limit is exceeded.
runWithLimit :: Replacer -> CodeUnitIndex -> Text -> Maybe Text
runWithLimit r (CodeUnitIndex maxLength) text = unsafePerformIO $
  withForeignPtr (replacerHandle r) $ \h -> withSlice text $ \hay ->
  alloca $ \outPtr -> alloca $ \outLen -> alloca $ \exceeded -> do
    let limit = if maxLength == maxBound then maxBound else fromIntegral maxLength
    _ <- amCall "am_replacer_run" [] $ c_am_replacer_run h (caseToC (replacerCase r)) hay limit outPtr outLen exceeded
    ex <- peek exceeded
    if ex /= 0 then pure Nothing else do
      p <- peek outPtr
      n <- fromIntegral <$> peek outLen
      -- copy the library's buffer into a fresh ByteArray#, then release it
      arr <- stToIO $ do { dst <- TextArray.new n; TextArray.copyFromPointer dst 0 p n; TextArray.unsafeFreeze dst }
      c_am_free p
      pure (Just (Text arr 0 n))
{-# NOINLINE runWithLimit #-}
