-- | Drop-in for "Data.Text.AhoCorasick.Searcher" (reference: src/Data/Text/AhoCorasick/Searcher.hs:14-27).
-- NOT COMPILED HERE -- see INTEGRATION.md.
module Data.Text.AhoCorasick.Searcher
  ( Searcher, build, buildWithValues, buildNeedleIdSearcher, containsAny, containsAll
  , needles, numNeedles, automaton, caseSensitivity, setCaseSensitivity, mapSearcher
  ) where

import Data.Hashable (Hashable)
import Data.Text.CaseSensitivity (CaseSensitivity (..))
import Data.Text.Utf8 (Text)
import Foreign.ForeignPtr (withForeignPtr)
import Foreign.Marshal (alloca)
import Foreign.Storable (peek)
import System.IO.Unsafe (unsafePerformIO)

import qualified Data.Text.AhoCorasick.Automaton as Aho
import Data.Text.AhoCorasick.FFI

data Searcher v = Searcher                         -- Searcher.hs:61-66
  { searcherCaseSensitive :: CaseSensitivity
  , searcherNeedles :: [(Text, v)]
  , searcherNumNeedles :: Int
  , searcherAutomaton :: Aho.AcMachine v
  }

build :: CaseSensitivity -> [Text] -> Searcher ()                                   -- :110-111
build case_ = buildWithValues case_ . fmap (\n -> (n, ()))

buildWithValues :: Hashable v => CaseSensitivity -> [(Text, v)] -> Searcher v      -- :115-118
buildWithValues case_ ns = Searcher case_ ns (length ns) (Aho.buildWithCase case_ ns)

buildNeedleIdSearcher :: CaseSensitivity -> [Text] -> Searcher Int                 -- :167-169
buildNeedleIdSearcher case_ ns = buildWithValues case_ (zip ns [0 ..])

needles :: Searcher v -> [(Text, v)]
needles = searcherNeedles
numNeedles :: Searcher v -> Int
numNeedles = searcherNumNeedles
automaton :: Searcher v -> Aho.AcMachine v
automaton = searcherAutomaton
caseSensitivity :: Searcher v -> CaseSensitivity
caseSensitivity = searcherCaseSensitive

setCaseSensitivity :: CaseSensitivity -> Searcher v -> Searcher v                   -- :142-145 (rebuilds the device image)
setCaseSensitivity case_ s = s { searcherCaseSensitive = case_, searcherAutomaton = Aho.buildWithCase case_ (searcherNeedles s) }

mapSearcher :: (a -> b) -> Searcher a -> Searcher b                                 -- :121-125 (payloads live on the host)
mapSearcher f s = s { searcherNeedles = fmap (fmap f) (searcherNeedles s), searcherAutomaton = fmap f (searcherAutomaton s) }

-- | `containsAny` (:156-164): one am_contains_any call (the kernel sets a flag; other CTAs stop at the next tile).
containsAny :: Searcher () -> Text -> Bool
containsAny s text = unsafePerformIO $ withForeignPtr (Aho.machineHandle (automaton s)) $ \h ->
  Aho.withSlice text $ \hay -> alloca $ \out -> do
    rc <- c_am_contains_any h hay out
    if rc /= amOk then Aho.amError "am_contains_any" else (/= 0) <$> peek out

-- | `containsAll` (:173-187).
containsAll :: Searcher Int -> Text -> Bool
containsAll s text = unsafePerformIO $ withForeignPtr (Aho.machineHandle (automaton s)) $ \h ->
  Aho.withSlice text $ \hay -> alloca $ \out -> do
    rc <- c_am_contains_all h hay out
    if rc /= amOk then Aho.amError "am_contains_all" else (/= 0) <$> peek out
